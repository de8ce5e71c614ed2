mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for k in fwd32 fwd64 fwd128 fwd32s; do timeout 60 python tools/run_kernel.py $k 20; done > gpurun_out/kernels.log 2>&1; cat gpurun_out/kernels.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
