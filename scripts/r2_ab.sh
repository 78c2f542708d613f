for v in "$@"; do
  IBK_LIB=$PWD/scripts/variants/$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sample-parity --e2e-steps 0 2>gpurun_out/ab_$v.err > gpurun_out/ab_$v.json
  python scripts/bench_brief.py < gpurun_out/ab_$v.json | sed "s/^/$v: /"; python -c "import json,sys; d=json.load(open('gpurun_out/ab_$v.json')); print('   check', d['check']['ok'], d['check']['adjointness_rel_err'])"
done
