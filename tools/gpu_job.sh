mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vnet.py -m gpu -q -x -k "conv1x1 or mri or bf16_tensor_core or f32_logits or fused_evaluation" 2>&1 | tail -4
timeout 300 python tools/bench_extra.py mri 2>&1 | tail -1
MSB_NO_PDL=1 timeout 300 python tools/step_timeline.py 2 mri > gpurun_out/timeline_mri.log 2>&1; grep -n "conv1x1\|dice_ce\|traced" gpurun_out/timeline_mri.log | cut -c1-130
