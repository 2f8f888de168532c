"""GPU: the engine called from several host threads at once (lx(1) runs fsm_determinise /
fsm_minimise on different fsm objects from a pthread pool, reference src/lx/main.c:338-475;
ctypes releases the GIL during the calls)."""
import os
import threading

import numpy as np
import pytest

import goldenio
import reflib
import libfsm_b200 as L
from libfsm_b200 import workloads
from test_oracle_determinise import assert_isomorphic

pytestmark = pytest.mark.gpu


def test_concurrent_determinise_minimise_exec(oracle):
    det = goldenio.load_det_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_determinise.npz"))
    mins = {c["name"]: c for c in goldenio.load_det_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_minimise.npz"))}
    work = [c for c in det if c["dfa"].nstates >= 5][:8] + [c for c in det if c["name"].startswith("cfg5:")]
    exec_case = next(c for c in goldenio.load_exec_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_exec.npz")) if c["name"] == "cfg2:ragged")
    errors, results = [], {}

    def worker(tid):
        try:
            for rnd in range(3):
                for k, c in enumerate(work):
                    if (k + rnd) % 4 != tid:
                        continue
                    d = L.determinise(c["nfa"])
                    m = L.minimise(d)
                    results[(tid, rnd, c["name"])] = (d, m)
                with L.Dfa(exec_case["fsm"]) as dfa:
                    got = dfa.exec_batch(exec_case["base"], exec_case["offsets"])
                    if not (got == exec_case["expect_amortised"]).all():
                        errors.append(f"thread {tid}: exec records differ")
        except Exception as e:                                   # noqa: BLE001
            errors.append(f"thread {tid}: {e!r}")

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    assert results
    for (tid, rnd, name), (d, m) in results.items():
        c = next(x for x in work if x["name"] == name)
        assert_isomorphic(oracle, d, c["dfa"])
        want_min = mins["min:" + name]["dfa"]
        if want_min.nstates:
            assert_isomorphic(oracle, m, want_min)
