mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 30 --warmup 5 --graph-ddp > gpurun_out/bench_n2_graph.log 2>&1; tail -3 gpurun_out/bench_n2_graph.log | cut -c1-400
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader
