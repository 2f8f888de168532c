#!/usr/bin/env python3
"""fsm_minimise of the config-5 generator's DFA: GPU (K3) vs the reference's fsm_minimise on this
box's CPU (oracle/_ref, single thread).  WORDS/LENGTH select the size."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import reflib, libfsm_b200 as L
from libfsm_b200 import workloads

words = int(os.environ.get("WORDS", 2000)); length = int(os.environ.get("LENGTH", 50))
nfa = workloads.config5_nfa(words, length, seed=12345)
dfa = L.determinise(nfa)
L.minimise(L.determinise(workloads.config5_nfa(20, 6)))
ts = []
for _ in range(3):
    t0 = time.perf_counter(); m = L.minimise(dfa); ts.append(time.perf_counter() - t0)
out = {"dfa_states_in": dfa.nstates, "dfa_states_min": m.nstates, "gpu_s": min(ts), "gpu_s_all": ts, "stats": L.minimise_stats()}
if reflib.have_ref():
    R = reflib.Ref()
    h = R.from_flat(nfa); R.determinise(h)
    t0 = time.perf_counter(); R.minimise(h); t1 = time.perf_counter()
    out["cpu_reference_s"] = t1 - t0; out["cpu_reference_states_min"] = R.countstates(h); out["speedup"] = (t1 - t0) / min(ts)
    R.free(h)
print(json.dumps(out))
