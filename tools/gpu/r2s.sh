# round 2, pass s: the whole GPU suite + smoke + the default bench line (+ config 4, 1) on the final tree
set -x
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
tail -8 gpurun_out/r2s_pytest.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/r2s_smoke.log 2>&1; tail -2 gpurun_out/r2s_smoke.log
timeout 600 python bench.py > gpurun_out/r2s_bench_default.json 2> gpurun_out/r2s_bench_default.err; cut -c1-200 gpurun_out/r2s_bench_default.json
timeout 400 python bench.py --config 4 > gpurun_out/r2s_bench_cfg4.json 2> gpurun_out/r2s_bench_cfg4.err; cut -c1-200 gpurun_out/r2s_bench_cfg4.json
timeout 300 python bench.py --config 1 > gpurun_out/r2s_bench_cfg1.json 2> gpurun_out/r2s_bench_cfg1.err; cut -c1-200 gpurun_out/r2s_bench_cfg1.json
DFAS=utf8: timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_launches_cfg4.csv python bench.py --config 4 --steps 2 --warmup 1 > /dev/null 2>&1
