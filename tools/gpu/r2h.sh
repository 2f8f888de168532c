# lines-kernel iteration: parity tests for every lines-kernel user, kernel-only numbers, bench config 3, ncu capture
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_config3.py tests/test_gpu_eager.py tests/test_gpu_exec.py -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
WHAT=cfg3 timeout 600 python tools/bench_r2.py > gpurun_out/r2h_micro.jsonl 2> gpurun_out/r2h_micro.err
timeout 600 python bench.py --config 3 > gpurun_out/r2h_bench.jsonl 2> gpurun_out/r2h_bench.err
WHAT=cfg3 NLINES=2000000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_lines -s 3 -c 1 -o gpurun_out/r2h_lines_eager python tools/bench_r2.py > gpurun_out/r2h_ncu.log 2>&1
tail -3 gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_micro.jsonl | cut -c1-400
