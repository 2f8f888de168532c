# multi-GPU pass: bash tools/gpu/r2g_multi.sh N [all]   (one node, torchrun as the driver launches it)
N=$1; MODE=${2:-short}
set -x
mkdir -p gpurun_out
OUT=gpurun_out/r2g_n$N.jsonl; ERR=gpurun_out/r2g_n$N.err
: > $OUT; : > $ERR
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" >> $OUT 2>> $ERR; echo "rc=$? $*" >> $ERR; }
if [ "$MODE" = cfg3 ]; then run --config 3; grep -c . $OUT; grep "rc=" $ERR; exit 0; fi
run --config 2
run --config 2 --gather-records full
run --config 4
run --config 3
if [ "$MODE" = all ]; then
  run --config 2 --dist adversarial
  run --config 2 --gather nccl
  run --config 2 --impl reference --steps 2 --warmup 1
fi
grep -c . $OUT; grep "rc=" $ERR
