# round 2, pass l: K1b fused small-automaton form (k1b_rep.cuh) -- parity, memcheck, A/B timing, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream_rep.py tests/test_gpu_stream.py -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -15 gpurun_out/r2l_pytest.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stream_rep.py -x -q -k "random_small and 7-0.001 or never_merge and 7-4" > gpurun_out/r2l_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2l_memcheck.log
tail -8 gpurun_out/r2l_memcheck.log
DFAS=utf8: timeout 300 python tools/bench_stream.py > gpurun_out/r2l_stream_rep.jsonl 2> gpurun_out/r2l_stream_rep.err
FSM_B200_STREAM_REP=0 DFAS=utf8: timeout 300 python tools/bench_stream.py > gpurun_out/r2l_stream_old.jsonl 2> gpurun_out/r2l_stream_old.err
cat gpurun_out/r2l_stream_rep.jsonl gpurun_out/r2l_stream_old.jsonl
NBYTES=$((1<<26)) DFAS=utf8: timeout 300 python tools/bench_stream.py >> gpurun_out/r2l_stream_rep.jsonl 2>> gpurun_out/r2l_stream_rep.err
timeout 300 python bench.py --config 1 > gpurun_out/r2l_bench_cfg1.json 2> gpurun_out/r2l_bench_cfg1.err
timeout 400 python bench.py --config 4 > gpurun_out/r2l_bench_cfg4.json 2> gpurun_out/r2l_bench_cfg4.err
tail -3 gpurun_out/r2l_stream_rep.jsonl; cat gpurun_out/r2l_bench_cfg1.json gpurun_out/r2l_bench_cfg4.json | cut -c1-600
DFAS=utf8: timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l_launches.csv python tools/bench_stream.py > /dev/null 2>&1
DFAS=utf8: timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1b_rep_kernel -s 2 -c 1 -f -o gpurun_out/r2l_rep python tools/bench_stream.py > gpurun_out/r2l_ncu.log 2>&1
ls -la gpurun_out | tail -20
