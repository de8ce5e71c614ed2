mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python tools/bench_extra.py mri > gpurun_out/mri.log 2>&1; tail -1 gpurun_out/mri.log
MSB_NO_PDL=1 timeout 300 python tools/step_timeline.py 2 mri > gpurun_out/timeline_mri.log 2>&1; head -14 gpurun_out/timeline_mri.log | cut -c1-130
