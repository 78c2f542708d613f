timeout 900 python -m pytest tests -m gpu -q -x -k "ranks_as_contexts" > gpurun_out/r2_t10.log 2>&1; tail -25 gpurun_out/r2_t10.log
