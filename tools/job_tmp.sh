OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for b in 32 8; do MSB_BUCKET_MB=$b MSB_NO_PDL=1 timeout 300 $TR --master-port 2980$b tools/step_timeline.py 3 > $OUT/r2d_timeline_n2_b$b.log 2>&1; echo "timeline b=$b rc=$?"; grep -A6 "NCCL report" $OUT/r2d_timeline_n2_b$b.log | cut -c1-600; done
for b in 8 16 64; do timeout 300 $TR --master-port 2981$b bench.py --gpus 2 --steps 20 --warmup 5 --bucket-mb $b > $OUT/r2d_bench_n2_b$b.log 2>&1; echo "n2 graph bucket=$b rc=$?"; tail -1 $OUT/r2d_bench_n2_b$b.log | cut -c1-200; done
