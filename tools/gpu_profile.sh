#!/bin/bash
# GPU-box job: one `ncu --set full` capture per kernel family (B200_PROFILING.md recipe; one launch each, source import
# on).  Usage: gpurun --timeout 1500 -- bash tools/gpu_profile.sh <tag> case:regex [case:regex ...]
# e.g.  wgrad32:conv_k5_wgrad  fwd64:conv_k5_fwd  splitk256:conv_k5_fwd|splitk_finalize  head20:eval_head
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for spec in "$@"; do
  case_=${spec%%:*}; regex=${spec#*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s 4 -c 2 -f \
    -o $OUT/${TAG}_${case_} python tools/run_kernel.py $case_ 3 > $OUT/${TAG}_${case_}_ncu.log 2>&1
  echo "$case_ rc=$?"
  # gpurun brings back at most 64 MiB: keep the text summary (+ the per-section details page), drop the report
  python tools/ncu_summary.py $OUT/${TAG}_${case_}.ncu-rep >> $OUT/${TAG}_ncu_summary.txt 2>&1
  ncu -i $OUT/${TAG}_${case_}.ncu-rep --page details 2>/dev/null | grep -E "^  [a-zA-Z]|Duration|Throughput|Registers|Shared Memory|Theoretical|Achieved|Stall|L2|DRAM|Tensor|Executed Ipc|Issue Slots|One or More Eligible|No Eligible" | head -120 > $OUT/${TAG}_${case_}_details.txt
  if [ "${KEEP_REP:-}" != "$case_" ]; then rm -f $OUT/${TAG}_${case_}.ncu-rep; fi
done
cat $OUT/${TAG}_ncu_summary.txt
