# round 2, pass m: k1b_rep with L2 prefetch-size hints / explicit prefetch, in-CTA fold -- parity, knob sweep, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream_rep.py tests/test_gpu_stream.py -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -5 gpurun_out/r2m_pytest.log
KNOBS="FSM_B200_REP_L2HINT=0;FSM_B200_REP_L2HINT=1;FSM_B200_REP_L2HINT=0,FSM_B200_REP_PREFETCH=512;FSM_B200_REP_L2HINT=0,FSM_B200_REP_PREFETCH=1024;FSM_B200_REP_L2HINT=0,FSM_B200_REP_PREFETCH=2048;FSM_B200_REP_L2HINT=1,FSM_B200_REP_PREFETCH=512;FSM_B200_REP_L2HINT=1,FSM_B200_REP_PREFETCH=1024;FSM_B200_REP_L2HINT=1,FSM_B200_REP_PREFETCH=2048;FSM_B200_REP_L2HINT=1,FSM_B200_REP_PREFETCH=4096" DFAS=utf8: timeout 300 python tools/bench_stream.py > gpurun_out/r2m_knobs.jsonl 2> gpurun_out/r2m_knobs.err
cat gpurun_out/r2m_knobs.jsonl
DFAS=utf8: timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1b_rep -s 4 -c 2 -f -o gpurun_out/r2m_rep_hint python tools/bench_stream.py > gpurun_out/r2m_ncu.log 2>&1
KNOBS="FSM_B200_REP_L2HINT=1,FSM_B200_REP_PREFETCH=1024" DFAS=utf8: timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1b_rep_kernel -s 2 -c 1 -f -o gpurun_out/r2m_rep_hint_pf python tools/bench_stream.py >> gpurun_out/r2m_ncu.log 2>&1
ls -la gpurun_out | tail
