for N in "$@"; do
IBK_BENCH_PHASES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_n$N.json 2> gpurun_out/r2_n$N.err
tail -1 gpurun_out/r2_n$N.json | python scripts/bench_brief.py; grep "phases\]" gpurun_out/r2_n$N.err | tail -13
done
