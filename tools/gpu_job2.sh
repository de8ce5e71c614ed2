mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.log 2>&1; tail -1 gpurun_out/bench_n2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
timeout 300 python bench.py --no-graph --no-cpu-baseline > gpurun_out/bench_eager.log 2>&1; tail -1 gpurun_out/bench_eager.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 eager', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
timeout 300 python tools/bench_extra.py infer > gpurun_out/infer.log 2>&1; tail -2 gpurun_out/infer.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/check_sync_bn_2gpu.py f32 2>&1 | grep -E "^rank"
