timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t7.log 2>&1; tail -3 gpurun_out/r2_t7.log
