# what the driver runs at round end, in one call: smoke(), the GPU test suite, the default bench line
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_gputests.log 2>&1; tail -2 gpurun_out/r02_gputests.log
( time python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err ) 2>&1 | grep real
python scripts/bench_brief.py < gpurun_out/r02_bench_default.json
