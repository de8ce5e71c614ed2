mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_transforms.py tests/test_gpu_vnet.py -m gpu -q -k "transform or rotation or flip or full_size or dataset or evaluation or eval_forward" > gpurun_out/pytest_tf.log 2>&1; tail -25 gpurun_out/pytest_tf.log
