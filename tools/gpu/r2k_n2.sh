set -x
mkdir -p gpurun_out
OUT=gpurun_out/r2k_n2.jsonl; ERR=gpurun_out/r2k_n2.err
: > $OUT; : > $ERR
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 "$@" >> $OUT 2>> $ERR; echo "rc=$? $*" >> $ERR; }
run --config 2
run --config 4 --e2e-steps 1
grep "rc=" $ERR
