mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.log 2>&1; tail -2 gpurun_out/bench_n2.log | cut -c1-700
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-300
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/bench_n1_nograph.log 2>&1; tail -1 gpurun_out/bench_n1_nograph.log | cut -c1-300
