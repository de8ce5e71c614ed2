OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/r2f_pytest.log 2>&1; echo "pytest rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/r2f_bench.log 2>&1; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/r2f_launches.csv python bench.py --no-graph --no-cpu-baseline --steps 2 --warmup 1 > $OUT/r2f_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
MSB_NO_PDL=1 timeout 300 python tools/step_timeline.py 3 > $OUT/r2f_timeline.log 2>&1; echo "timeline rc=$?"
( for v in "MSB_X=0" "MSB_DEBUG6=32" "MSB_DEBUG0=330" "MSB_DEBUG0=380" "MSB_DEBUG6=16"; do echo "== wgrad32 $v"; env $v timeout 120 python tools/run_kernel.py wgrad32 10; done
  echo "== mriwgrad256 default / clustered(8)"; timeout 120 python tools/run_kernel.py mriwgrad256 5; MSB_DEBUG6=8 timeout 120 python tools/run_kernel.py mriwgrad256 5
  echo "== mriwgrad128 default / clustered(8)"; timeout 120 python tools/run_kernel.py mriwgrad128 5; MSB_DEBUG6=8 timeout 120 python tools/run_kernel.py mriwgrad128 5
  for c in wgrad64 wgrad32s k2scatter k2scatter_acc k2gather; do timeout 100 python tools/run_kernel.py $c 10; done ) 2>&1 | grep -v Warn > $OUT/r2f_kernels.log
KEEP_REP=wgrad32 bash tools/gpu_profile.sh r2f "wgrad32:conv_k5_wgrad2" "wgrad64:conv_k5_wgrad2" "fwd64:conv_k5_fwd" "fwd128:conv_k5_fwd" "splitk256:conv_k5_fwd|splitk_finalize" "wgrad128:conv_k5_wgrad_kernel" "mriwgrad256:conv_k5_wgrad_kernel" "mriwgrad128:conv_k5_wgrad_kernel" "head20:eval_head" "loss20:dice_ce" "trilinear:trilinear" "preprocess:resample" "k2scatter_acc:conv_k2s2" "k2gather:conv_k2s2" "k2scatter:conv_k2s2" > $OUT/r2f_profile_stdout.log 2>&1
export PYTEST_ADDOPTS="-v"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "layout or bn_prelu or in_tr or dice_ce or conv1x1 or momentum or trilinear or preprocess_matches or down_conv_and_up_conv" > $OUT/r2f_racecheck.log 2>&1; echo "racecheck rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "not full_size" > $OUT/r2f_memcheck.log 2>&1; echo "memcheck rc=$?"
ls -la $OUT | head -40; du -sh $OUT
echo ==== PYTEST; tail -8 $OUT/r2f_pytest.log
echo ==== BENCH; tail -1 $OUT/r2f_bench.log | cut -c1-220
echo ==== KERNELS; cat $OUT/r2f_kernels.log
echo ==== SAN; grep -E "SUMMARY|passed|failed" $OUT/r2f_racecheck.log $OUT/r2f_memcheck.log | tail -6
