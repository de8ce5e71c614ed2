timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "strided_tensor_core or down_conv_and_up_conv or k2s2" 2>&1 | tail -2
for w in k2gather k2scatter k2scatter_acc; do timeout 60 python tools/run_kernel.py $w 20 2>&1 | tail -1; done
