"""GPU: K2 with FSM_B200_DET_REFERENCE_NUMBERING -- the DFA must be the reference's `struct fsm`
STATE FOR STATE (no canonicalisation): same numbering, same per-state groups in the same order,
same end bits and end-id sets, against the DFAs the reference recorded in
tests/golden/golden_determinise.npz.

The numbering functions (libfsm_b200/csrc/refnum.h) are also verified bit-exactly on the CPU
(tests/test_refnum_host.py: goldens + live reference on random NFAs).  Passed on the driver's B200
at the end of round 1; plain (strict) tests since round 2.
"""
import os

import numpy as np
import pytest

import goldenio
import libfsm_b200 as L
from libfsm_b200 import workloads

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

CASES = goldenio.load_det_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_determinise.npz"))
CASES = [c for c in CASES if c["nfa"].hasstart]


def assert_same_fsm(got, want):
    assert got.nstates == want.nstates
    assert got.start == want.start == 0
    assert np.array_equal(np.asarray(got.is_end).astype(bool), np.asarray(want.is_end).astype(bool))
    assert np.array_equal(got.group_off, want.group_off)
    G = int(want.group_off[-1])
    assert np.array_equal(np.asarray(got.group_to)[:G], np.asarray(want.group_to)[:G])
    assert np.array_equal(np.asarray(got.group_symbols).reshape(-1, 4)[:G], np.asarray(want.group_symbols).reshape(-1, 4)[:G])
    for s in range(want.nstates):
        assert list(got.endids_of(s)) == list(want.endids_of(s))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_reference_numbering_matches_recorded_reference(case):
    got = L.determinise(case["nfa"], numbering="reference")
    assert_same_fsm(got, case["dfa"])
    assert L.determinise_stats()["ms_numbering"] > 0


def test_reference_numbering_is_a_renumbering_of_the_default(oracle):
    """Full-size property (BASELINE config 5): both numberings give the same canonical DFA."""
    from test_oracle_determinise import assert_isomorphic
    nfa = workloads.config5_nfa()
    a = L.determinise(nfa, numbering="bfs")
    b = L.determinise(nfa, numbering="reference")
    assert a.nstates == b.nstates
    assert_isomorphic(oracle, a, b)


def test_fsm_cli_prints_the_reference_text_byte_for_byte(tmp_path):
    """`fsm -pd` of the relinked fsm(1) (fsm_determinise -> K2) with
    FSM_B200_DET_NUMBERING=reference prints exactly what the reference's fsm(1) prints."""
    import subprocess
    from test_gpu_shim import FSM_B200, FSM_REF, to_fsm5
    if not (os.path.exists(FSM_B200) and os.path.exists(FSM_REF)):
        pytest.skip("relinked fsm(1) not built")
    env = dict(os.environ, FSM_B200_DET_NUMBERING="reference")
    ran = 0
    for c in CASES:
        txt = to_fsm5(c["nfa"])
        if txt is None or c["dfa"].nstates > 2000:
            continue
        inp = tmp_path / "in.fsm"
        inp.write_text(txt)
        got = subprocess.run([FSM_B200, "-pd"], stdin=open(inp), capture_output=True, timeout=120, env=env)
        want = subprocess.run([FSM_REF, "-pd"], stdin=open(inp), capture_output=True, timeout=120)
        assert got.returncode == 0 and want.returncode == 0, (c["name"], got.stderr)
        assert got.stdout == want.stdout, c["name"]
        ran += 1
    assert ran >= 10
