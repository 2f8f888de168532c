# round 2, pass p: the whole GPU suite on the tree with the fused K1b form, then bench.py configs 4 and 1
set -x
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -8 gpurun_out/r2p_pytest.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/r2p_smoke.log 2>&1; tail -2 gpurun_out/r2p_smoke.log
timeout 400 python bench.py --config 4 > gpurun_out/r2p_bench_cfg4.json 2> gpurun_out/r2p_bench_cfg4.err; cut -c1-200 gpurun_out/r2p_bench_cfg4.json
timeout 300 python bench.py --config 1 > gpurun_out/r2p_bench_cfg1.json 2> gpurun_out/r2p_bench_cfg1.err; cut -c1-200 gpurun_out/r2p_bench_cfg1.json
