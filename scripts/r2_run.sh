timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t18.log 2>&1; tail -5 gpurun_out/r2_t18.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sample-parity --e2e-steps 1 2>/dev/null > gpurun_out/r2_b18.json; python scripts/bench_brief.py < gpurun_out/r2_b18.json | cut -c1-250; python -c "
import json; d=json.load(open('gpurun_out/r2_b18.json')); print(d['gpu_launches'], d['check']['ok'])"
