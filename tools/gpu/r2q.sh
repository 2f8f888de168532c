# round 2, pass q: k1b_rep_tma (input by 2-D TMA tiles, compact per-lane table) -- parity, A/B against the 256-bit-load form, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream_rep.py tests/test_gpu_stream.py -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -6 gpurun_out/r2q_pytest.log
FSM_B200_REP_TMA_STAGES=2 timeout 600 python -m pytest tests/test_gpu_stream_rep.py -x -q > gpurun_out/r2q_pytest_st2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest_st2.log
tail -2 gpurun_out/r2q_pytest_st2.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stream_rep.py -x -q -k "random_small and 7-0.001 or never_merge and 7-4 or death_offsets and 0-4000" > gpurun_out/r2q_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2q_memcheck.log
tail -4 gpurun_out/r2q_memcheck.log
KNOBS="FSM_B200_REP_TMA=0;FSM_B200_REP_TMA=1;FSM_B200_REP_TMA=1,FSM_B200_REP_TMA_STAGES=3;FSM_B200_REP_TMA=1,FSM_B200_REP_TMA_STAGES=2" DFAS=utf8: timeout 300 python tools/bench_stream.py > gpurun_out/r2q_knobs.jsonl 2> gpurun_out/r2q_knobs.err
cat gpurun_out/r2q_knobs.jsonl
KNOBS="FSM_B200_REP_TMA=1" DFAS=utf8: timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1b_rep_tma -s 2 -c 1 -f -o gpurun_out/r2q_rep_tma python tools/bench_stream.py > gpurun_out/r2q_ncu.log 2>&1
tail -3 gpurun_out/r2q_ncu.log
