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
  echo "$case_ rc=$?"; tail -2 $OUT/${TAG}_${case_}_ncu.log
done
