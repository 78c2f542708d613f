#!/bin/bash
# Opt-in paths that were written or only partly measured at the end of round 1: run under gpurun,
#   gpurun --timeout 900 -- 'bash scripts/pending_experiments.sh'
# 1. the TMA reducing store of the spread block (4.08 ms against 4.28 ms): full parity suite, then the bench
# 2. the 2x2x2 cluster spread with the reducing store of the share (never run on a GPU)
# 3. the brick-colour spread with 320 threads per CTA (never run on a GPU)
set -u
brief() { python scripts/bench_brief.py; }
echo "== IBK_SPREAD_REDUCE=1: full GPU suite"; IBK_SPREAD_REDUCE=1 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== IBK_SPREAD_REDUCE=1: bench"; IBK_SPREAD_REDUCE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | brief
echo "== IBK_SPREAD_CLUSTER=2: parity"; IBK_SPREAD_CLUSTER=2 timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -3
echo "== IBK_SPREAD_CLUSTER=2: bench"; IBK_SPREAD_CLUSTER=2 timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | brief
echo "== IBK_SPREAD_WIDE=1 (320 threads, 100-marker windows): parity + bench"; IBK_SPREAD_WIDE=1 timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -3
IBK_SPREAD_WIDE=1 timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | brief
echo "== default: bench"; python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | brief
