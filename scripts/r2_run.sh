timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t23.log 2>&1; tail -4 gpurun_out/r2_t23.log
