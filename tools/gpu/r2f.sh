# round-2 GPU pass 2: tests added since pass 1, ncu capture of the tail-scheduled tile kernel, stream benches
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_stream.py tests/test_gpu_shim.py tests/test_gpu_exec.py tests/test_gpu_determinise.py -m gpu -x -q -rs > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2f_smoke.log
WHAT=cfg2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_krange -s 3 -c 1 -o gpurun_out/r2f_krange_tail python tools/bench_r2.py > gpurun_out/r2f_ncu1.log 2>&1
: > gpurun_out/r2f_bench.jsonl
for c in 4 1; do
  timeout 600 python bench.py --config $c >> gpurun_out/r2f_bench.jsonl 2>> gpurun_out/r2f_bench.err
done
tail -5 gpurun_out/r2f_pytest.log; tail -3 gpurun_out/r2f_smoke.log
