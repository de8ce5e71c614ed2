#!/bin/bash
# GPU-box job: the driver's round-end sequence in one call (pytest -m gpu, smoke, bench lines of every BASELINE config,
# reference arm, ncu launch list of the bench command).  Usage: gpurun --timeout 2400 -- bash tools/gpu_ci.sh [tag]
# Everything lands in gpurun_out/ (scratch); summaries worth keeping are copied to profiles/ by hand.
TAG=${1:-ci}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 300 python -m pytest tests/test_gpu_boundary.py -q -s -k loader -p no:cacheprovider 2>&1 | grep "reader_cost per log window" | tee $OUT/${TAG}_loader.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.log 2>&1; echo "bench rc=$?"; tail -1 $OUT/${TAG}_bench.log | cut -c1-400
for cfg in mri_bf16 vnet128_fp32ddp preprocess; do
  timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 > $OUT/${TAG}_bench_${cfg}.log 2>&1
  echo "bench $cfg rc=$?"; tail -1 $OUT/${TAG}_bench_${cfg}.log | cut -c1-300
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 $OUT/${TAG}_bench_ref.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --no-graph --no-cpu-baseline --steps 2 --warmup 1 > $OUT/${TAG}_bench_under_ncu.log 2>&1
echo "ncu launch list rc=$?"
