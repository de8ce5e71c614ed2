OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_fullsize.py tests/test_gpu_kernels.py -q --timeout 600 -p no:cacheprovider -k "data_parallel or loader or checkpoints or fullsize or benchmark_shape or mri_512 or split_k or wgrad or down_conv or strided" > $OUT/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/r2c_pytest.log
timeout 300 $TR --master-port 29701 bench.py --gpus 2 --check > $OUT/r2c_check.log 2>&1; echo "check rc=$?"; grep -v Warning $OUT/r2c_check.log | tail -12
timeout 300 $TR --master-port 29702 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/r2c_bench_n2_graph.log 2>&1; echo "n2 graph rc=$?"; tail -1 $OUT/r2c_bench_n2_graph.log | cut -c1-250
timeout 300 $TR --master-port 29703 bench.py --gpus 2 --steps 20 --warmup 5 --no-graph > $OUT/r2c_bench_n2_eager.log 2>&1; echo "n2 eager rc=$?"; tail -1 $OUT/r2c_bench_n2_eager.log | cut -c1-250
for c in 4 8 16; do NCCL_MAX_CTAS=$c timeout 300 $TR --master-port 2971$c bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/r2c_bench_n2_graph_cta$c.log 2>&1; echo "n2 graph maxctas=$c rc=$?"; tail -1 $OUT/r2c_bench_n2_graph_cta$c.log | cut -c1-200; done
timeout 300 $TR --master-port 29704 bench.py --gpus 2 --steps 20 --warmup 5 --sync-bn > $OUT/r2c_bench_n2_syncbn.log 2>&1; echo "n2 syncbn rc=$?"; tail -1 $OUT/r2c_bench_n2_syncbn.log | cut -c1-250
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r2c_bench_n1.log 2>&1; tail -1 $OUT/r2c_bench_n1.log | cut -c1-200
( timeout 100 python tools/run_kernel.py wgrad32 10; timeout 100 python tools/run_kernel.py wgrad32s 10; for c in k2scatter k2scatter_acc k2gather; do timeout 100 python tools/run_kernel.py $c 10; done ) 2>&1 | grep -v Warn
