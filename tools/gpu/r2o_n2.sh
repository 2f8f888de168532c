# round 2, pass o (2 GPUs): config 4 with the shard maps left on the device + one NCCL all-gather
set -x
mkdir -p gpurun_out
OUT=gpurun_out/r2o_n2.jsonl; ERR=gpurun_out/r2o_n2.err
: > $OUT; : > $ERR
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 "$@" >> $OUT 2>> $ERR; echo "rc=$? $*" >> $ERR; }
run --config 4 --e2e-steps 1
grep "rc=" $ERR; tail -5 $ERR; cut -c1-300 $OUT
