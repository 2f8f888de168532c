#!/usr/bin/env python3
"""Round-2 kernel-only measurements (inputs resident in HBM, CUDA events, inputs larger than L2):
  cfg2   BASELINE config 2, both distributions: k-stride with ALU classification (RNG), k-stride with
         class LUTs (FSM_B200_KSTRIDE_LUT=1), LANE
  cfg3   BASELINE config 3 (tests/golden/golden_cfg3.npz, built by the reference): the 128-pattern
         eager-output DFA over NLINES ragged lines (lines kernel, records + id bitsets), the same
         automaton through the plain entry point, and the start-anchored end-id variant
One JSON line per measurement.  WHAT=cfg2,cfg3 selects."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import goldenio  # noqa: E402
import libfsm_b200 as L  # noqa: E402
from libfsm_b200 import workloads  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


def cfg2():
    n, length = int(os.environ.get("N", 1 << 20)), 1024
    cases = {c["name"]: c for c in goldenio.load_exec_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_exec.npz"))}
    fsm = cases["cfg2:uniform"]["fsm"]
    for adversarial in (False, True):
        dev = workloads.cfg2_device(n, length, adversarial, seed=42)
        with L.Dfa(fsm) as dfa:
            out = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
            ref = None
            for name, variant, env in (("kstride-alu", "kstride", {}), ("kstride-lut", "kstride", {"FSM_B200_KSTRIDE_LUT": "1"}), ("lane", "lane", {})):
                for blk in os.environ.get("BLOCKS", "1024").split(","):
                    os.environ.update(env); os.environ["FSM_B200_KSTRIDE_BLOCK"] = blk
                    L.set_exec_variant(variant)
                    ms, mn = timed(lambda: dfa.exec_batch(dev, stride=length, length=length, n=n, out=out))
                    for k in env:
                        del os.environ[k]
                    if ref is None:
                        ref = out.clone()
                    print(json.dumps({"what": "cfg2", "kernel": name, "block": int(blk), "adversarial": adversarial, "ms": round(ms, 4), "ms_min": round(mn, 4),
                                      "GBps": round(n * length / ms / 1e6, 1), "frac_hbm": round(n * length / ms / 1e6 / PEAK, 4),
                                      "agree": bool(torch.equal(out, ref)), "krange": dfa.info["krange"]}), flush=True)
                    if variant != "kstride":
                        break
        del dev


def cfg3():
    g = goldenio.load_cfg3()
    nlines = int(os.environ.get("NLINES", 10_000_000))
    _, inst = workloads.cfg3_patterns()
    _, ainst = workloads.cfg3_anchored_patterns()
    for name, fsm, instances, at_start in (("eager128", g["eager"]["fsm"], inst, False), ("anchored128", g["anchored"]["fsm"], ainst, True)):
        base, offsets = workloads.cfg3_lines_device(nlines, instances, seed=7, at_start=at_start)
        total = int(offsets[-1])
        with L.Dfa(fsm) as dfa:
            info = dfa.info
            out = torch.empty((nlines, 16), dtype=torch.uint8, device="cuda")
            for blk in os.environ.get("LINES_BLOCKS", "768").split(","):
                os.environ["FSM_B200_LINES_BLOCK"] = blk
                if info["eager_ids"]:
                    ms, mn = timed(lambda: dfa.exec_batch_eager(base, offsets))
                    rec, masks = dfa.exec_batch_eager(base, offsets)
                    torch.cuda.synchronize()
                    fired = float((masks != 0).any(dim=1).float().mean())
                    print(json.dumps({"what": "cfg3", "dfa": name, "entry": "exec_batch_eager (records + id bitsets)", "block": int(blk), "lines": nlines, "bytes": total,
                                      "states": fsm.nstates, "lines_blob_bytes": info["lines_blob_bytes"], "ms": round(ms, 3), "ms_min": round(mn, 3),
                                      "GBps_all_bytes_read": round(total / ms / 1e6, 1), "frac_hbm": round(total / ms / 1e6 / PEAK, 4),
                                      "lines_with_ids": round(fired, 4), "match_rate": round(float((L.results_from_torch(rec)["ret"] == 1).mean()), 4)}), flush=True)
                for skip in ("0", "1"):
                    if skip == "1":
                        os.environ["FSM_B200_NO_ABSORB_SKIP"] = "1"
                    ms, mn = timed(lambda: dfa.exec_batch(base, offsets, out=out))
                    os.environ.pop("FSM_B200_NO_ABSORB_SKIP", None)
                    rec = L.results_from_torch(out)
                    read = int(rec["consumed"].sum()) if skip == "1" else None
                    print(json.dumps({"what": "cfg3", "dfa": name, "entry": "exec_batch (records)", "absorb_exit": skip == "0", "block": int(blk), "lines": nlines, "bytes": total,
                                      "states": fsm.nstates, "ms": round(ms, 3), "GBps_bytes_covered": round(total / ms / 1e6, 1),
                                      "bytes_walked": read, "GBps_bytes_walked": None if read is None else round(read / ms / 1e6, 1),
                                      "frac_hbm_covered": round(total / ms / 1e6 / PEAK, 4), "match_rate": round(float((rec["ret"] == 1).mean()), 4)}), flush=True)
        del base, offsets


if __name__ == "__main__":
    for w in os.environ.get("WHAT", "cfg2,cfg3").split(","):
        {"cfg2": cfg2, "cfg3": cfg3}[w]()
