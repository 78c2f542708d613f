timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t28.log 2>&1; tail -3 gpurun_out/r2_t28.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b28.json 2> gpurun_out/r2_b28.err; python scripts/bench_brief.py < gpurun_out/r2_b28.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_c3_launches.csv python bench.py --config C3 --steps 1 --warmup 1 --no-cpu-baseline --no-sample-parity --e2e-steps 0 > gpurun_out/r2_c3_ncu.log 2>&1
