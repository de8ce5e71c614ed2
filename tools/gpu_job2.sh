mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_sync_bn.py f32 > gpurun_out/sync_bn.log 2>&1; grep -E "rank|Error|error" gpurun_out/sync_bn.log | head
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_sync_bn.py bf16 >> gpurun_out/sync_bn.log 2>&1; grep -E "rank|Error|error" gpurun_out/sync_bn.log | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.log 2>&1; tail -1 gpurun_out/bench_n2.log | cut -c1-300
