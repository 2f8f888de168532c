"""CPU, world_size 2 over gloo: the multi-GPU host logic (range sharding by bytes, one
all-gather of result records, stream-map composition).  The per-shard "scan" here is done by
the oracle -- this test is about the plumbing around the kernels, not the kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import goldenio
import reflib
from libfsm_b200 import sharding, workloads
from libfsm_b200.desc import RESULT_DTYPE

CASES = {c["name"]: c for c in goldenio.load_exec_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_exec.npz"))}


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = reflib.Oracle()
        fsm = CASES["cfg2:uniform"]["fsm"]
        base, offsets = workloads.ragged_lines_host(4001, 0, 200, seed=11, alphabet=b"a" * 25 + bytes(range(0x20, 0x7F)))
        ranges = sharding.shard_ranges_by_bytes(offsets, world)
        lo, hi = ranges[rank]
        mine = oracle.exec_batch(fsm, base, offsets[lo:hi + 1])
        # one all-gather of fixed-size records: pad every shard to the largest
        cap = max(h - l for l, h in ranges)
        buf = np.zeros(cap, dtype=RESULT_DTYPE); buf[:hi - lo] = mine
        send = torch.from_numpy(buf.view(np.uint8).copy())
        recv = torch.empty(world * send.numel(), dtype=torch.uint8)
        dist.all_gather_into_tensor(recv, send)
        parts = recv.numpy().view(RESULT_DTYPE).reshape(world, cap)
        full = np.concatenate([parts[r][:h - l] for r, (l, h) in enumerate(ranges)])
        want = oracle.exec_batch(fsm, base, offsets)
        ok_batch = bool((full == want).all())

        # stream: per-shard maps from the oracle, all-gather, compose
        utf = next(c for n, c in CASES.items() if n.startswith("utf8:"))["fsm"]
        tab = oracle.flatten(utf)
        S = utf.nstates
        dead = S
        stream = workloads.utf8_host(50000, seed=3)
        stream = np.concatenate([stream, np.frombuffer(b"\xff", dtype=np.uint8), stream[:100]]) if os.environ.get("BAD") else stream
        br = sharding.byte_ranges(stream.size, world, align=16)
        blo, bhi = br[rank]
        ms = np.zeros(S + 1, np.uint32); md = np.full(S + 1, sharding.NO_DEAD, np.uint64); mf = np.full(S + 1, 0xFFFFFFFF, np.uint32)
        for s in range(S + 1):
            if s == dead:
                ms[s] = dead; md[s] = 0; mf[s] = dead
                continue
            st = s
            for k in range(blo, bhi):
                nx = tab[st, stream[k]]
                if nx == 0xFFFFFFFF:
                    md[s] = k - blo; mf[s] = st; st = dead
                    break
                st = nx
            ms[s] = st
        rec = torch.from_numpy(np.concatenate([ms.astype(np.uint64), md, mf.astype(np.uint64)]).astype(np.int64))
        allrec = torch.empty(world * rec.numel(), dtype=torch.int64)
        dist.all_gather_into_tensor(allrec, rec)
        allrec = allrec.numpy().astype(np.uint64).reshape(world, 3, S + 1)
        st, consumed, died = sharding.compose_stream_maps(
            utf.start, dead, [h - l for l, h in br], allrec[:, 0].astype(np.uint32), allrec[:, 1], allrec[:, 2].astype(np.uint32))
        ret, end, cons = oracle.exec(utf, stream.tobytes())
        ok_stream = (cons == consumed) and (end == st) and ((ret == 1) == (not died and bool(utf.is_end[st])))
        # the same exchange in the record layout bench.py --config 4 gathers (fsm_b200_stream_map_entry:
        # state | died << 32, dead offset; [nstates] records left on the device by exec_stream_map_dev_async)
        from libfsm_b200.engine import stream_map_arrays
        dflag = md[:S] != sharding.NO_DEAD
        rec2 = np.zeros((S, 2), dtype=np.uint64)
        rec2[:, 0] = np.where(dflag, mf[:S], ms[:S]).astype(np.uint64) | (dflag.astype(np.uint64) << np.uint64(32))
        rec2[:, 1] = md[:S]
        all2 = torch.empty((world * S, 2), dtype=torch.int64)          # gloo wants the concatenated shape (NCCL also takes the stacked one)
        dist.all_gather_into_tensor(all2, torch.from_numpy(rec2.view(np.int64)))
        S2, D2, F2 = stream_map_arrays(all2.numpy().reshape(world, S, 2))
        st2, consumed2, died2 = sharding.compose_stream_maps(utf.start, dead, [h - l for l, h in br], list(S2), list(D2), list(F2))
        ok_stream = ok_stream and (st2, consumed2, died2) == (st, consumed, died)
        q.put((rank, ok_batch, ok_stream))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bad", [False, True])
def test_world2_gloo(bad):
    if bad:
        os.environ["BAD"] = "1"
    else:
        os.environ.pop("BAD", None)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    os.environ.pop("BAD", None)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res


def test_shard_ranges_cover_and_balance():
    _, offsets = workloads.ragged_lines_host(10000, 0, 300, seed=2)
    for world in (1, 2, 3, 8):
        rs = sharding.shard_ranges_by_bytes(offsets, world)
        assert rs[0][0] == 0 and rs[-1][1] == 10000
        assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
        sizes = [int(offsets[h] - offsets[l]) for l, h in rs]
        assert max(sizes) - min(sizes) <= 300 * 2
    assert sharding.byte_ranges(1000, 3) == [(0, 320), (320, 656), (656, 1000)]


def test_stream_map_records_decode_and_compose():
    """fsm_b200_stream_map_entry records (state u32, died u32, dead_off u64) as gathered by bench.py --config 4:
    decoded into the arrays compose_stream_maps folds (host logic only, no device)."""
    from libfsm_b200.engine import stream_map_arrays
    rec = np.zeros((2, 3, 2), dtype=np.uint64)                  # 2 shards, 3 entry states
    rec[0, :, 0] = [1, 2, 0]; rec[0, :, 1] = 0xFFFFFFFFFFFFFFFF   # shard 0: 0->1, 1->2, 2->0, nobody dies
    rec[1, 0, 0] = 2; rec[1, 0, 1] = 0xFFFFFFFFFFFFFFFF
    rec[1, 1, 0] = 0 | (1 << 32); rec[1, 1, 1] = 7               # shard 1 from state 1: dies at offset 7, read in state 0
    rec[1, 2, 0] = 1; rec[1, 2, 1] = 0xFFFFFFFFFFFFFFFF
    S, D, F = stream_map_arrays(rec.view(np.int64))
    assert S.dtype == np.uint32 and D.dtype == np.uint64
    assert sharding.compose_stream_maps(0, 3, [100, 50], list(S), list(D), list(F)) == (0, 107, True)
    assert sharding.compose_stream_maps(2, 3, [100, 50], list(S), list(D), list(F)) == (2, 150, False)
