timeout 600 python -m pytest tests -m gpu -q -x -k "walls" > gpurun_out/r2_t20.log 2>&1; tail -15 gpurun_out/r2_t20.log
