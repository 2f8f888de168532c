"""GPU: K3 -- fsm_b200_minimise against the reference's fsm_minimise outputs (golden, pipeline
order) and, at BASELINE config 5's scale, against the oracle; canonical-form isomorphism."""
import os

import numpy as np
import pytest

import goldenio
import reflib
import libfsm_b200 as L
from libfsm_b200 import workloads
from test_oracle_determinise import assert_isomorphic

pytestmark = pytest.mark.gpu

CASES = goldenio.load_det_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_minimise.npz"))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_minimise_isomorphic_to_reference(oracle, case):
    got = L.minimise(case["nfa"])
    want = case["dfa"]
    if want.nstates == 0:
        assert got.nstates == 0
        return
    assert_isomorphic(oracle, got, want)
    assert oracle.isdfa(got)
    assert L.minimise(got).nstates == got.nstates            # idempotent
    assert L.minimise_stats()["kernel_launches"] > 0


def test_not_a_dfa_is_refused():
    nfa = L.FlatFsm.from_edges(3, 0, [2], [(0, ord("a"), 1), (0, ord("a"), 2)])
    with pytest.raises(L.FsmB200Error):
        L.minimise(nfa)


def test_nothing_can_match_gives_empty_fsm():
    f = L.FlatFsm.from_edges(3, 0, [], [(0, ord("a"), 1), (1, ord("b"), 2)])     # no end state
    assert L.minimise(f).nstates == 0


def test_determinise_then_minimise_then_exec(oracle):
    """The re(1) pipeline on the engine alone: determinise (K2) -> minimise (K3) -> exec (K1)
    == exec of the reference's determinised+minimised DFA on the same inputs."""
    det = goldenio.load_det_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_determinise.npz"))
    case = next(c for c in det if c["name"] == "epsunion:9pats")
    mine = L.minimise(L.determinise(case["nfa"]))
    ref_min = next(c for c in CASES if c["name"] == "min:epsunion:9pats")["dfa"]
    assert_isomorphic(oracle, mine, ref_min)
    rng = np.random.default_rng(4)
    al = np.frombuffer(b"abcdxy0123fooranch", dtype=np.uint8)
    strs = [al[rng.integers(0, al.size, int(rng.integers(0, 12)))].tobytes() for _ in range(2000)] + [b"abc", b"foobar", b"12x", b""]
    base, off = reflib.offsets_for(strs)
    want = oracle.exec_batch(ref_min, base, off)
    with L.Dfa(mine) as dfa:
        got = dfa.exec_batch(base, off)
    assert (got["ret"] == want["ret"]).all() and (got["consumed"] == want["consumed"]).all()


def test_config5_scale(oracle):
    """Minimise the 48.5k-state DFA of the config-5 generator (1000 words x 50): isomorphic to
    the oracle's minimal DFA (the oracle is pinned to the reference on the golden cases)."""
    nfa = workloads.config5_nfa(1000, 50, seed=12345)
    dfa = L.determinise(nfa)
    got = L.minimise(dfa)
    print("minimise stats:", L.minimise_stats())
    want = oracle.minimise(dfa)
    assert got.nstates == want.nstates
    assert_isomorphic(oracle, got, want)
