timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t12.log 2>&1; tail -15 gpurun_out/r2_t12.log
