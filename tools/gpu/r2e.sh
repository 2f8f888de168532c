# round-2 GPU pass: full -m gpu suite, bench lines for every config, tail-schedule A/B, ncu launch list
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
: > gpurun_out/r2e_bench.jsonl
for t in 1 0; do for d in uniform adversarial; do
  FSM_B200_KRTILE_TAIL=$t timeout 300 python bench.py --config 2 --dist $d >> gpurun_out/r2e_bench.jsonl 2>> gpurun_out/r2e_bench.err
done; done
for c in 1 3 4 5; do
  timeout 600 python bench.py --config $c >> gpurun_out/r2e_bench.jsonl 2>> gpurun_out/r2e_bench.err
done
timeout 600 python bench.py --impl reference --config 2 --steps 3 --warmup 1 >> gpurun_out/r2e_bench.jsonl 2>> gpurun_out/r2e_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches_bench.csv python bench.py --steps 4 --warmup 3 > gpurun_out/r2e_bench_under_ncu.log 2>&1
tail -5 gpurun_out/r2e_pytest.log
