# round-2 evidence (one GPU): GPU test suite, bench line, reference arm, launch list, full-set captures of the two dominant kernels
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_gputests.log 2>&1; tail -2 gpurun_out/r02_gputests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_reference.json 2> gpurun_out/r02_reference.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sample-parity --e2e-steps 1 > gpurun_out/r02_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spread_march -s 1 -c 1 -f -o gpurun_out/r02_spread_march python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sample-parity --e2e-steps 0 > gpurun_out/r02_ncu_s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:interp_rot -s 1 -c 1 -f -o gpurun_out/r02_interp_rot python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sample-parity --e2e-steps 0 > gpurun_out/r02_ncu_i.log 2>&1
python scripts/bench_brief.py < gpurun_out/r02_bench.json; tail -c 400 gpurun_out/r02_reference.json
