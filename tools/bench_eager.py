#!/usr/bin/env python3
"""Eager outputs: throughput of k1_eager.cu -- the batch kernel that also
returns the bitset of eager-output ids fired per input -- next to the plain K1 kernels on the same
DFA and inputs, device-resident, CUDA events.  DFA: the reference's own construction recorded in
tests/golden/golden_eager.npz (fsm_union_repeated_pattern_group of unanchored patterns ->
determinise -> minimise).  Checks the fired sets of a sample against the oracle first."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import goldenio, reflib, libfsm_b200 as L

case = next(c for c in goldenio.load_eager_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_eager.npz"))
            if c["name"] == os.environ.get("CASE", "group:overlap"))
fsm = case["min"]
n, length = int(os.environ.get("N", 1 << 20)), int(os.environ.get("LEN", 256))
rng = np.random.default_rng(3)
host = rng.choice(np.frombuffer(b"abxyz", dtype=np.uint8), size=(n, length))
dev = torch.from_numpy(host).cuda()
oracle = reflib.Oracle()
out = {"case": case["name"], "dfa_states": fsm.nstates, "n": n, "len": length}
with L.Dfa(fsm) as dfa:
    out["eager_ids"] = [int(x) for x in dfa.eager_ids()]
    rec, masks = dfa.exec_batch_eager(dev, stride=length, length=length, n=n)
    torch.cuda.synchronize()
    m = masks.cpu().numpy().view(np.uint64)
    for i in range(0, n, max(1, n // 200)):
        want = oracle.exec_eager(fsm, host[i].tobytes())[1]
        assert dfa.fired_ids(m[i]) == want, i
    def timed(fn, reps=20):
        for _ in range(3): fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    ms_e = timed(lambda: dfa.exec_batch_eager(dev, stride=length, length=length, n=n))
    ms_p = timed(lambda: dfa.exec_batch(dev, stride=length, length=length, n=n))
    out.update(ms_eager=ms_e, gbs_eager=n * length / ms_e / 1e6, ms_plain=ms_p, gbs_plain=n * length / ms_p / 1e6)
print(json.dumps(out))
