timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t27.log 2>&1; tail -3 gpurun_out/r2_t27.log
IBK_BENCH_PHASES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2_n2.json 2> gpurun_out/r2_n2.err
tail -1 gpurun_out/r2_n2.json | python scripts/bench_brief.py; grep "phases\]" gpurun_out/r2_n2.err | tail -14
