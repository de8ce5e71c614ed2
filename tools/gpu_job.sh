mkdir -p gpurun_out
timeout 600 python tools/bench_extra.py fp32 > gpurun_out/fp32.log 2>&1; tail -1 gpurun_out/fp32.log
timeout 300 python tools/bench_extra.py mri > gpurun_out/mri.log 2>&1; tail -1 gpurun_out/mri.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "dice_ce" 2>&1 | tail -2
