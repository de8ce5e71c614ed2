mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vnet.py -m gpu -q -k "trilinear or deepsup" > gpurun_out/pytest_ds.log 2>&1; tail -30 gpurun_out/pytest_ds.log
timeout 300 python tools/bench_extra.py augment > gpurun_out/augment.log 2>&1; tail -3 gpurun_out/augment.log
