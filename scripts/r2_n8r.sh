for R in 8 16; do
IBK_COMM_RESERVE_SMS=$R timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$R bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2_n8_r$R.json 2> gpurun_out/r2_n8_r$R.err
echo "reserve $R:"; tail -1 gpurun_out/r2_n8_r$R.json | python scripts/bench_brief.py
done
