# N=8: is the compact (4-byte) gather payload slower than full records, or was it a first-run effect?
set -x
mkdir -p gpurun_out
OUT=gpurun_out/r2i_n8.jsonl; ERR=gpurun_out/r2i_n8.err
: > $OUT; : > $ERR
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 8 "$@" >> $OUT 2>> $ERR; echo "rc=$? $*" >> $ERR; }
run --config 2 --gather-records full --e2e-steps 2
run --config 2 --e2e-steps 2
run --config 2 --e2e-steps 2
run --config 2 --gather-records full --e2e-steps 2
run --config 2 --no-consumer --e2e-steps 2
grep "rc=" $ERR
