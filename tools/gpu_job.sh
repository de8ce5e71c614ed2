mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 100 python tools/run_kernel.py bn32 20 > gpurun_out/bn.log 2>&1; cat gpurun_out/bn.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-300
MSB_NO_PDL=1 timeout 200 python tools/step_timeline.py 3 > gpurun_out/timeline_nopdl.log 2>&1; grep -E "traced|bn_" gpurun_out/timeline_nopdl.log | cut -c1-130
