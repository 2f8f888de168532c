#!/usr/bin/env python3
"""K1b throughput: one long input resident in HBM (config 4's per-GPU share and a config-2 DFA stream)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import goldenio, libfsm_b200 as L
from libfsm_b200 import workloads

cases = goldenio.load_exec_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_exec.npz"))
def case(p): return next(c for c in cases if c["name"].startswith(p))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
nbytes = int(os.environ.get("NBYTES", 1 << 31))
for name in os.environ.get("DFAS", "utf8:,cfg2:uniform,cfg1:digits").split(","):
    fsm = case(name)["fsm"]
    if name == "utf8:":
        block = workloads.utf8_host(1 << 24, seed=6)
        block = np.concatenate([block, np.full((-block.size) % 16, ord("a"), dtype=np.uint8)])
        dev = torch.from_numpy(block).cuda().repeat(nbytes // block.size)
    else:
        dev = workloads.cfg2_device(nbytes // 1024, 1024, False, seed=1).reshape(-1)
    with L.Dfa(fsm) as dfa:
        # KNOBS: ';'-separated settings, each a ','-separated list of ENV=value pairs applied for that measurement
        for knobs in os.environ.get("KNOBS", "").split(";"):
            pairs = [kv.split("=", 1) for kv in knobs.split(",") if "=" in kv]
            for k, v in pairs: os.environ[k] = v
            for chunk in os.environ.get("CHUNKS", "default").split(","):
                if chunk == "default": os.environ.pop("FSM_B200_STREAM_CHUNK", None)
                else: os.environ["FSM_B200_STREAM_CHUNK"] = chunk
                for _ in range(3): r = dfa.exec_stream(dev)
                torch.cuda.synchronize()
                ts = []
                for _ in range(int(os.environ.get("REPS", 9))):
                    t0 = time.perf_counter(); r = dfa.exec_stream(dev); ts.append(time.perf_counter() - t0)
                ms = float(np.median(ts)) * 1e3
                gbs = dev.numel() / ms / 1e6
                print(json.dumps({"dfa": name, "states": fsm.nstates, "bytes": int(dev.numel()), "chunk": chunk, "knobs": knobs, "ms": round(ms, 4),
                                  "GBps": round(gbs, 1), "frac_hbm": round(gbs / peak, 4), "result": r}), flush=True)
            for k, _ in pairs: os.environ.pop(k, None)
    del dev
