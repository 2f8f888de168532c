#!/usr/bin/env python3
"""Kernel-only throughput of every K1 variant on BASELINE config 2 (inputs resident in HBM,
CUDA events, inputs larger than L2).  Prints one JSON line per (variant, distribution, pad)."""
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


def main():
    n = int(os.environ.get("N", 1 << 20)); length = 1024
    cases = {c["name"]: c for c in goldenio.load_exec_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_exec.npz"))}
    fsm = cases["cfg2:uniform"]["fsm"]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    variants = os.environ.get("VARIANTS", "kstride,lane,tile64,tile32,tile128,tile64x3").split(",")
    pads = os.environ.get("PADS", "4,0").split(",")
    for adversarial in (False, True):
        dev = workloads.cfg2_device(n, length, adversarial, seed=42)
        for pad in pads:
            os.environ["FSM_B200_ROW_PAD"] = pad
            with L.Dfa(fsm) as dfa:
                ref_out = None
                for v in variants:
                    L.set_exec_variant(v)
                    out = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
                    for _ in range(3):
                        dfa.exec_batch(dev, stride=length, length=length, n=n, out=out)
                    torch.cuda.synchronize()
                    if ref_out is None:
                        ref_out = out.clone()
                    ok = bool(torch.equal(out, ref_out))
                    times = []
                    for _ in range(10):
                        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                        e0.record(); dfa.exec_batch(dev, stride=length, length=length, n=n, out=out); e1.record()
                        torch.cuda.synchronize(); times.append(e0.elapsed_time(e1))
                    ms = float(np.median(times))
                    gbs = n * length / ms / 1e6
                    print(json.dumps({"variant": v, "adversarial": adversarial, "row_pad": int(pad), "ms": round(ms, 4),
                                      "ms_min": round(min(times), 4), "GBps": round(gbs, 1),
                                      "frac_hbm": round(gbs / peaks["hbm_gbs"], 4), "agree": ok}), flush=True)
        del dev


if __name__ == "__main__":
    main()
