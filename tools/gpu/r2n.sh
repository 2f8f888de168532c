# round 2, pass n: k1b_rep with 2 / 3 / 4 sector buffers (1..3 loads in flight per lane) x L2 hint; lines kernel L2 hint A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream_rep.py tests/test_gpu_stream.py -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -4 gpurun_out/r2n_pytest.log
FSM_B200_REP_NBUF=4 timeout 600 python -m pytest tests/test_gpu_stream_rep.py -x -q > gpurun_out/r2n_pytest_nbuf4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest_nbuf4.log
FSM_B200_REP_NBUF=2 FSM_B200_REP_L2HINT=0 timeout 600 python -m pytest tests/test_gpu_stream_rep.py -x -q > gpurun_out/r2n_pytest_nbuf2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest_nbuf2.log
tail -2 gpurun_out/r2n_pytest_nbuf4.log gpurun_out/r2n_pytest_nbuf2.log
KNOBS="FSM_B200_REP_NBUF=2,FSM_B200_REP_L2HINT=1;FSM_B200_REP_NBUF=3,FSM_B200_REP_L2HINT=0;FSM_B200_REP_NBUF=3,FSM_B200_REP_L2HINT=1;FSM_B200_REP_NBUF=4,FSM_B200_REP_L2HINT=0;FSM_B200_REP_NBUF=4,FSM_B200_REP_L2HINT=1" DFAS=utf8: timeout 300 python tools/bench_stream.py > gpurun_out/r2n_knobs.jsonl 2> gpurun_out/r2n_knobs.err
cat gpurun_out/r2n_knobs.jsonl
KNOBS="FSM_B200_REP_NBUF=3,FSM_B200_REP_L2HINT=1" DFAS=utf8: timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1b_rep_kernel -s 2 -c 1 -f -o gpurun_out/r2n_rep_nbuf3 python tools/bench_stream.py > gpurun_out/r2n_ncu.log 2>&1
KNOBS="FSM_B200_REP_NBUF=4,FSM_B200_REP_L2HINT=1" DFAS=utf8: timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1b_rep_kernel -s 2 -c 1 -f -o gpurun_out/r2n_rep_nbuf4 python tools/bench_stream.py >> gpurun_out/r2n_ncu.log 2>&1
WHAT=cfg3 NLINES=3000000 timeout 300 python tools/bench_r2.py > gpurun_out/r2n_lines_hint0.jsonl 2> gpurun_out/r2n_lines.err
FSM_B200_LINES_L2HINT=1 WHAT=cfg3 NLINES=3000000 timeout 300 python tools/bench_r2.py > gpurun_out/r2n_lines_hint1.jsonl 2>> gpurun_out/r2n_lines.err
cut -c1-400 gpurun_out/r2n_lines_hint0.jsonl gpurun_out/r2n_lines_hint1.jsonl
