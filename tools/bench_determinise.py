#!/usr/bin/env python3
"""BASELINE config 5: fsm_determinise of the 100 001-state synthetic NFA, GPU (K2) vs the
reference's fsm_determinise on this box's CPU (oracle/_ref, single thread: the reference is
single-threaded)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import reflib, libfsm_b200 as L
from libfsm_b200 import workloads

words = int(os.environ.get("WORDS", 2000)); length = int(os.environ.get("LENGTH", 50))
numbering = os.environ.get("NUMBERING") or None      # None (library default) | bfs | reference
nfa = workloads.config5_nfa(words, length, seed=12345)
L.determinise(workloads.config5_nfa(50, 10), numbering=numbering)   # warm up: context, module load
ts = []
for _ in range(3):
    t0 = time.perf_counter(); dfa = L.determinise(nfa, numbering=numbering); ts.append(time.perf_counter() - t0)
st = L.determinise_stats()
edges = dfa.nstates * 256
out = {"numbering": numbering or "default", "nfa_states": nfa.nstates, "dfa_states": dfa.nstates, "dfa_groups": int(dfa.group_to.size),
       "gpu_s": min(ts), "gpu_s_all": ts, "stats": st, "dfa_edges_per_s_gpu": edges / min(ts)}
if reflib.have_ref():
    R = reflib.Ref()
    h = R.from_flat(nfa)
    t0 = time.perf_counter(); R.determinise(h); t1 = time.perf_counter()
    out["cpu_reference_s"] = t1 - t0
    out["cpu_reference_dfa_states"] = R.countstates(h)
    out["speedup"] = (t1 - t0) / min(ts)
    if numbering == "reference":
        import numpy as np
        ref_dfa = R.flatten(h)
        out["identical_to_reference"] = bool(
            ref_dfa.nstates == dfa.nstates and np.array_equal(ref_dfa.group_off, dfa.group_off)
            and np.array_equal(ref_dfa.group_to, dfa.group_to[:ref_dfa.group_to.size])
            and np.array_equal(np.asarray(ref_dfa.is_end).astype(bool), np.asarray(dfa.is_end).astype(bool)))
    R.free(h)
print(json.dumps(out))
