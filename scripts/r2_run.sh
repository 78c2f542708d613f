timeout 600 python -m pytest tests -m gpu -q -x -k "binning" > gpurun_out/r2_t19.log 2>&1; tail -12 gpurun_out/r2_t19.log
