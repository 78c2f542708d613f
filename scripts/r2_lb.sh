ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_lb_launches.csv python scripts/loopback_step.py 512 23 1 > gpurun_out/r2_lb.log 2>&1
tail -2 gpurun_out/r2_lb.log
timeout 300 python scripts/loopback_step.py 512 23 5 2>&1 | tail -1
