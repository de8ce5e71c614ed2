"""medicalseg.utils.env_util.get_sys_env (reference env_util/sys_env.py:81-120): the environment table train.py prints."""
import platform
import sys


def get_sys_env():
    import torch
    env = {"platform": platform.platform(), "Python": sys.version.replace("\n", ""),
           "PyTorch": torch.__version__, "CUDA (torch)": torch.version.cuda}
    gpu = torch.cuda.is_available()
    env["GPUs used"] = torch.cuda.device_count() if gpu else 0
    if gpu:
        env["GPU"] = ["GPU %d: %s" % (i, torch.cuda.get_device_name(i)) for i in range(torch.cuda.device_count())]
    try:
        from medicalseg_b200 import _lib
        env["libmedseg_b200"] = "v%d (%s)" % (_lib.call("msb_version"), _lib.LIB_PATH)
    except Exception as e:  # the table is informational: a missing library is reported, the model classes raise
        env["libmedseg_b200"] = "NOT LOADED: %s" % e
    return env
