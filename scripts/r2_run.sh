timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t22.log 2>&1; tail -5 gpurun_out/r2_t22.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sample-parity --e2e-steps 2 2>/dev/null > gpurun_out/r2_b22.json; python scripts/bench_brief.py < gpurun_out/r2_b22.json | cut -c1-250
timeout 300 python bench.py --config C3 --markers uniform --steps 10 --warmup 3 --no-cpu-baseline --no-sample-parity --e2e-steps 1 2>/dev/null | python scripts/bench_brief.py | cut -c1-250
