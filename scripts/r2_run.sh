timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t25.log 2>&1; tail -4 gpurun_out/r2_t25.log
timeout 300 python scripts/loopback_step.py 512 23 3 2>&1 | tail -1
