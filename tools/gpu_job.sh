mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "strided_tensor_core" > gpurun_out/pytest_k2.log 2>&1; tail -2 gpurun_out/pytest_k2.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_k5_fwd_kernel -s 3 -c 1 -o gpurun_out/fwd32_final -f python tools/run_kernel.py fwd32 > gpurun_out/ncu_fwd32.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_k5_wgrad2_kernel -s 3 -c 1 -o gpurun_out/wgrad32_final -f python tools/run_kernel.py wgrad32 > gpurun_out/ncu_wgrad32.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 python tools/bench_extra.py preprocess > gpurun_out/preprocess.log 2>&1; tail -1 gpurun_out/preprocess.log
for k in fwd32 fwd64 fwd128 fwd256 fwd256s fwd128s wgrad32 wgrad64 wgrad128 wgrad256 k2scatter k2scatter_acc k2gather bn32; do timeout 60 python tools/run_kernel.py $k 20; done > gpurun_out/kernels.log 2>&1; cat gpurun_out/kernels.log
