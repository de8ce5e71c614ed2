mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vnet.py tests/test_gpu_kernels.py -m gpu -q -k "deepsup or trilinear" 2>&1 | tail -2
timeout 300 python tools/bench_extra.py deepsup > gpurun_out/deepsup.log 2>&1; tail -1 gpurun_out/deepsup.log
MSB_NO_PDL=1 timeout 300 python tools/step_timeline.py 2 deepsup > gpurun_out/timeline_deepsup.log 2>&1; grep -n "trilinear\|traced" gpurun_out/timeline_deepsup.log | cut -c1-130
