#!/bin/bash
# GPU-box job: compute-sanitizer over the kernel tests (memcheck / racecheck per kernel test).  The kernels
# under test use mbarrier rings, TMA (multicast) and red.global workspaces with a zero-on-exit contract.
# Usage: gpurun --timeout 1800 -- bash tools/gpu_sanitize.sh [tag] [pytest -k expression]
TAG=${1:-san}
KEXPR=${2:-"layout or bn_prelu or in_tr or down_conv or tcgen05 or split_k or evaluation_epilogue or tap_major or k2s2 or strided_tensor or dice_ce or conv1x1 or momentum or trilinear"}
OUT=gpurun_out
mkdir -p $OUT
export MSB_TEST_SMALL=1
for tool in memcheck racecheck; do
  timeout 800 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "$KEXPR" > $OUT/${TAG}_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a $OUT/${TAG}_${tool}.log
  grep -E "ERROR SUMMARY|passed|failed|Error" $OUT/${TAG}_${tool}.log | tail -5
done
