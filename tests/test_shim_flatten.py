"""CPU: the shim's struct fsm -> flat description walk (fsm_b200_flatten in
libfsm_b200/shim/fsm_b200_shim.c) against the test harness's independent walk, on automata
built by the reference's re_comp (NFAs with epsilons) and by its determinise+minimise (DFAs
with end ids).  Uses build/shim/flatten_dump (the flattener compiled over the unmodified
reference library; no GPU)."""
import os
import subprocess

import numpy as np
import pytest

import reflib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMP = os.path.join(ROOT, "build", "shim", "flatten_dump")

pytestmark = pytest.mark.skipif(not os.path.exists(DUMP), reason="flatten_dump not built (needs the reference tree at build time)")


def parse_dump(text):
    d = {}
    for line in text.strip().splitlines():
        k, *v = line.split()
        d[k] = v
    hdr = {"nstates": int(d["nstates"][0]), "start": int(d["nstates"][2]), "hasstart": int(d["nstates"][4])}
    groups = [g.split(":") for g in d.get("groups", [])]
    return hdr, {
        "is_end": np.array(d.get("is_end", []), dtype=np.uint8),
        "group_off": np.array(d["group_off"], dtype=np.uint64),
        "group_to": np.array([int(g[0]) for g in groups], dtype=np.uint32),
        "group_symbols": np.array([[int(x, 16) for x in g[1:]] for g in groups], dtype=np.uint64).reshape(-1, 4),
        "eps_off": np.array(d["eps_off"], dtype=np.uint64), "eps_to": np.array(d.get("eps_to", []), dtype=np.uint32),
        "endid_off": np.array(d["endid_off"], dtype=np.uint64), "endids": np.array(d.get("endids", []), dtype=np.uint32),
    }


@pytest.mark.parametrize("pattern", [r"ab*c|abd", r"a[ -~]{7}\z", r"^(GET|POST) /[a-z]+", r"[0-9]+\.[0-9]+", r"(foo|bar)+baz", "x{2,4}y"])
@pytest.mark.parametrize("det", [False, True])
def test_flatten_matches_harness(ref, pattern, det):
    p = subprocess.run([DUMP, pattern] + (["d"] if det else []), capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stderr
    hdr, got = parse_dump(p.stdout)
    h = ref.re_comp(pattern)
    if det:
        ref.determinise(h); ref.minimise(h); ref.setendid(h, 7)
    want = ref.flatten(h)
    ref.free(h)
    assert hdr == {"nstates": want.nstates, "start": want.start, "hasstart": int(want.hasstart)}
    for k, v in got.items():
        w = getattr(want, k)
        assert v.shape == w.shape and (v == w).all(), k
    if not det:
        assert int(want.eps_off[-1]) > 0          # the NFA really has epsilons
    else:
        assert int(want.endid_off[-1]) > 0 and (want.endids == 7).all()
