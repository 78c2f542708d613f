timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t9.log 2>&1; tail -5 gpurun_out/r2_t9.log
