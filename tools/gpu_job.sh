mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "three_pass or k5_tcgen05_forward or tap_major" > gpurun_out/pytest_a.log 2>&1; tail -3 gpurun_out/pytest_a.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_vnet.py -m gpu -q -x -k "f32x3" > gpurun_out/pytest_b.log 2>&1; tail -12 gpurun_out/pytest_b.log | cut -c1-300
timeout 300 python tools/bench_extra.py fp32x3 > gpurun_out/fp32x3.log 2>&1; tail -2 gpurun_out/fp32x3.log | cut -c1-300
