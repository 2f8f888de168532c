#!/usr/bin/env python3
"""BASELINE config 3 shape: an rx(1)-style union DFA of many PCRE patterns (per pattern
re_comp -> determinise -> minimise -> setendid(index); fsm_union_array; determinise, no final
minimise -- reference src/rx/main.c:487-566,1353,1371) scanned over ragged synthetic log lines
(length ~U[64,256]) resident in HBM.  Patterns are start-anchored templates so that the union
DFA stays buildable (unanchored unions explode combinatorially: 15 unanchored patterns already
gave 1.7 M states with the reference).  Needs the compiled reference (oracle/_ref) to BUILD the
DFA; parity of the scan is checked against the oracle on a sample."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import reflib, libfsm_b200 as L

NPAT = int(os.environ.get("NPAT", 128)); NLINES = int(os.environ.get("NLINES", 10_000_000)); SEED = 7
rng = np.random.default_rng(SEED)
words = ["ERROR", "WARN", "INFO", "DEBUG", "FATAL", "TRACE", "kernel", "sshd", "nginx", "cron", "systemd", "postfix", "docker", "kubelet"]
templates = [lambda w, k: f"^{w} [0-9]{{{k}}} ", lambda w, k: f"^{w}: user=[a-z]+ id=[0-9]{{{k}}}", lambda w, k: f"^{w}\\[[0-9]+\\]: ",
             lambda w, k: f"^[0-9]{{{k}}}\\.[0-9]+\\.[0-9]+\\.[0-9]+ {w}", lambda w, k: f"^{w} (GET|POST|PUT) /[a-z/]+ ", lambda w, k: f"^{w} [A-Z]{{{k}}}-[0-9]+"]
patterns, prefixes = [], []
while len(patterns) < NPAT:
    w = words[int(rng.integers(len(words)))] + str(int(rng.integers(0, 40))); k = int(rng.integers(1, 5)); t = int(rng.integers(len(templates)))
    p = templates[t](w, k)
    if p in patterns: continue
    patterns.append(p)
    ex = {0: f"{w} {'7' * k} ", 1: f"{w}: user=bob id={'4' * k}", 2: f"{w}[123]: ", 3: f"{'1' * k}.2.3.4 {w}", 4: f"{w} GET /a/b ", 5: f"{w} {'Q' * k}-99"}[t]
    prefixes.append(ex.encode())
R = reflib.Ref()
t0 = time.perf_counter(); h = R.union_dfa(patterns, state_limit=500000); t_build = time.perf_counter() - t0
fsm = R.flatten(h)
O = reflib.Oracle()

# lines: half start with a pattern instance, half are noise
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(SEED)
lens = torch.randint(64, 257, (NLINES,), device=dev, generator=g, dtype=torch.int64)
offsets = torch.zeros(NLINES + 1, dtype=torch.int64, device=dev); offsets[1:] = torch.cumsum(lens, 0)
total = int(offsets[-1])
base = torch.randint(0x20, 0x7F, (total,), dtype=torch.uint8, device=dev, generator=g)
pl = max(len(p) for p in prefixes)
ptab = torch.zeros((len(prefixes), pl), dtype=torch.uint8); plen = torch.zeros(len(prefixes), dtype=torch.int64)
for i, p in enumerate(prefixes):
    ptab[i, :len(p)] = torch.tensor(list(p), dtype=torch.uint8); plen[i] = len(p)
ptab, plen = ptab.to(dev), plen.to(dev)
pid = torch.randint(0, len(prefixes), (NLINES,), device=dev, generator=g)
use = torch.rand(NLINES, device=dev, generator=g) < 0.5
for k in range(pl):
    m = use & (plen[pid] > k)
    base[offsets[:-1][m] + k] = ptab[pid[m], k]

with L.Dfa(fsm) as dfa:
    out = torch.empty((NLINES, 16), dtype=torch.uint8, device=dev)
    for _ in range(3): dfa.exec_batch(base, offsets, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dfa.exec_batch(base, offsets, out=out); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    # parity on a sample of lines (bit-exact records + end-id sets)
    ns = 200000
    hb = base[:int(offsets[ns])].cpu().numpy(); ho = offsets[:ns + 1].cpu().numpy().astype(np.uint64)
    want = O.exec_batch(fsm, hb, ho, nthreads=min(32, os.cpu_count() or 1))
    got = L.results_from_torch(out[:ns])
    ok = bool((got == want).all())
    t0 = time.perf_counter(); refrec = R.exec_batch(h, hb, ho, mode=1, nthreads=os.cpu_count() or 1); t_cpu = time.perf_counter() - t0
    ok_ref = bool((refrec == got).all())
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    print(json.dumps({"workload": f"config 3 shape: {NPAT}-pattern rx-style union DFA, {NLINES} ragged lines", "dfa_states": fsm.nstates,
                      "table": dfa.info, "build_s_reference_cpu": t_build, "input_bytes": total, "kernel_ms": ms, "GBps": total / ms / 1e6,
                      "frac_hbm": total / ms / 1e6 / peak, "match_rate": float((got["ret"] == 1).mean()),
                      "bit_exact_vs_oracle_sample": ok, "bit_exact_vs_reference_sample": ok_ref,
                      "cpu_reference_amortised_GBps": int(ho[-1]) / t_cpu / 1e9, "cpu_threads": os.cpu_count()}))
