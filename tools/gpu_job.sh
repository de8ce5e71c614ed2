mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for k in wgrad32 wgrad64 wgrad128 wgrad256; do timeout 60 python tools/run_kernel.py $k 20; done > gpurun_out/kernels.log 2>&1; cat gpurun_out/kernels.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-300
timeout 200 python tools/step_timeline.py 3 > gpurun_out/timeline.log 2>&1; head -40 gpurun_out/timeline.log
