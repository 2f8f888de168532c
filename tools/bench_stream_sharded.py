#!/usr/bin/env python3
"""BASELINE config 4 shape: one long UTF-8 stream range-sharded over N GPUs (one process per
GPU, torchrun).  Every rank maps its byte range (K1b shard form), ONE all-gather of the [T]
(state, dead offset, dead-from) records, ordered composition on every rank.
  python -m torch.distributed.run --nproc-per-node N tools/bench_stream_sharded.py"""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import goldenio, reflib, libfsm_b200 as L
from libfsm_b200 import sharding, workloads

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
shard_bytes = int(os.environ.get("SHARD_BYTES", 1 << 31))
cases = goldenio.load_exec_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_exec.npz"))
fsm = next(c for c in cases if c["name"].startswith("utf8:"))["fsm"]
block = workloads.utf8_host(1 << 24, seed=6 + rank)
block = np.concatenate([block, np.full((-block.size) % 16, ord("a"), dtype=np.uint8)])
buf = torch.from_numpy(block).to(dev).repeat(shard_bytes // block.size)
n = int(buf.numel())
dfa = L.Dfa(fsm, device=local)
T = dfa.info["ntable_states"]
dead_row = None if dfa.info["complete"] else T - 1

def step():
    ms, md, mf = dfa.exec_stream_map(buf)
    rec = torch.from_numpy(np.concatenate([ms.astype(np.int64), md.view(np.int64), mf.astype(np.int64), np.array([n], np.int64)])).to(dev)
    if world > 1:
        allrec = torch.empty(world * rec.numel(), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allrec, rec)
    else:
        allrec = rec
    a = allrec.cpu().numpy().reshape(world, 3 * T + 1)
    lens = [int(x) for x in a[:, 3 * T]]
    return sharding.compose_stream_maps(fsm.start, dead_row, lens, a[:, :T].astype(np.uint32), a[:, T:2 * T].view(np.uint64) if False else a[:, T:2 * T].astype(np.uint64), a[:, 2 * T:3 * T].astype(np.uint32)), lens

for _ in range(2):
    (st, consumed, died), lens = step()
total = sum(lens)
assert not died and consumed == total and fsm.is_end[st], (st, consumed, died)
torch.cuda.synchronize()
if world > 1: dist.barrier()
ts = []
for _ in range(5):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
t = torch.tensor([float(np.median(ts))], dtype=torch.float64, device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
# corrupt one byte on the last rank: the composed verdict must carry its global offset
pos = n // 2 + 3
if rank == world - 1:
    old = int(buf[pos]); buf[pos] = 0xFF
(st2, consumed2, died2), _ = step()
want = sum(lens[:-1]) + pos
ok = died2 and want - 4 < consumed2 <= want
if rank == world - 1:
    buf[pos] = old
if rank == 0:
    sec = float(t.item())
    print(json.dumps({"workload": "config 4: UTF-8 validator, range-sharded stream", "n_gpus": world, "bytes_total": total,
                      "s_per_pass": sec, "GBps": total / sec / 1e9, "accepted": True, "first_invalid_offset_ok": bool(ok),
                      "states": fsm.nstates}))
assert ok, (consumed2, want)
dfa.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
