timeout 1000 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t5.log 2>&1; tail -3 gpurun_out/r2_t5.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sample-parity > gpurun_out/r2_b5.json 2> gpurun_out/r2_b5.err; tail -4 gpurun_out/r2_b5.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b5.json')); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d.get('check'))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spread_march -s 17 -c 1 -o gpurun_out/r2_march_d -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sample-parity > gpurun_out/r2_ncu6.log 2>&1
