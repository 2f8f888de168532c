# round 2, pass v (2 GPUs): the default bench line (config 2) on the final tree, as the driver's scaling run launches it
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/r2v_n2.jsonl 2> gpurun_out/r2v_n2.err; echo "rc=$?" >> gpurun_out/r2v_n2.err
tail -2 gpurun_out/r2v_n2.err; cut -c1-260 gpurun_out/r2v_n2.jsonl
