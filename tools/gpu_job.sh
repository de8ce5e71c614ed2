mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-400
for k in wgrad64 wgrad32; do timeout 60 python tools/run_kernel.py $k 20; done
timeout 300 python tools/bench_extra.py mri > gpurun_out/mri.log 2>&1; tail -1 gpurun_out/mri.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
