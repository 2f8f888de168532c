# round 2, pass t: pinned host pool for determinise / minimise results -- whole GPU suite, then bench.py --config 5 (and 4)
set -x
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
tail -6 gpurun_out/r2t_pytest.log
timeout 600 python bench.py --config 5 > gpurun_out/r2t_bench_cfg5.json 2> gpurun_out/r2t_bench_cfg5.err; cut -c1-300 gpurun_out/r2t_bench_cfg5.json
timeout 400 python bench.py --config 4 > gpurun_out/r2t_bench_cfg4.json 2> gpurun_out/r2t_bench_cfg4.err; cut -c1-200 gpurun_out/r2t_bench_cfg4.json
