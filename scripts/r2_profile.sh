# round-2 evidence: bench line, launch list, full-set captures of the two dominant kernels (one GPU)
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sample-parity --e2e-steps 1 > gpurun_out/r02_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spread_march -s 1 -c 1 -f -o gpurun_out/r02_spread_march python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sample-parity --e2e-steps 0 > gpurun_out/r02_ncu_s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:interp_rot -s 1 -c 1 -f -o gpurun_out/r02_interp_rot python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sample-parity --e2e-steps 0 > gpurun_out/r02_ncu_i.log 2>&1
tail -c 600 gpurun_out/r02_bench.json
