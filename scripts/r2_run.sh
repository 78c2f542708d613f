timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t26.log 2>&1; tail -3 gpurun_out/r2_t26.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b26.json 2> gpurun_out/r2_b26.err; python scripts/bench_brief.py < gpurun_out/r2_b26.json
