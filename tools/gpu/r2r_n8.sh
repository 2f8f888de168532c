# round 2, pass r (8 GPUs): config 4 (16 GiB range-sharded UTF-8 stream) with the fused K1b form + device-resident shard maps
set -x
mkdir -p gpurun_out
OUT=gpurun_out/r2r_n8.jsonl; ERR=gpurun_out/r2r_n8.err
: > $OUT; : > $ERR
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 8 "$@" >> $OUT 2>> $ERR; echo "rc=$? $*" >> $ERR; }
run --config 4 --e2e-steps 1
grep "rc=" $ERR; cut -c1-300 $OUT
