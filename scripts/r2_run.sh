timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 --e2e-steps 1 > gpurun_out/r2_b8.json 2> gpurun_out/r2_b8.err; python scripts/bench_brief.py < gpurun_out/r2_b8.json | cut -c1-300; tail -3 gpurun_out/r2_b8.err | cut -c1-300; python -c "
import json
for l in open('gpurun_out/r2_b8.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['gpu_launches'], d['check'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 10 --warmup 3 --e2e-steps 1 2>/dev/null | python scripts/bench_brief.py | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 10 --warmup 3 --e2e-steps 1 2>/dev/null | python scripts/bench_brief.py | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sample-parity --e2e-steps 1 2>/dev/null | python scripts/bench_brief.py | cut -c1-250
timeout 900 python -m pytest tests -m mgpu -q -x 2>&1 | tail -3
