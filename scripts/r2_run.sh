timeout 1000 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t3.log 2>&1; tail -5 gpurun_out/r2_t3.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b3.json 2> gpurun_out/r2_b3.err; tail -4 gpurun_out/r2_b3.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b3.json')); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'])"
