mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/bench_n8.log 2>&1; tail -1 gpurun_out/bench_n8.log | cut -c1-330
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/bench_n4.log 2>&1; tail -1 gpurun_out/bench_n4.log | cut -c1-330
