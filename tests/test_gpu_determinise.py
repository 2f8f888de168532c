"""GPU: K2 -- fsm_b200_determinise against the reference's fsm_determinise outputs (golden
fixtures incl. the reference's own tests/determinise + tests/eclosure inputs) and, at
BASELINE config 5's full size, against the oracle; DFAs compared in canonical form."""
import os

import numpy as np
import pytest

import goldenio
import reflib
import libfsm_b200 as L
from libfsm_b200 import workloads
from test_oracle_determinise import assert_isomorphic

pytestmark = pytest.mark.gpu

CASES = goldenio.load_det_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_determinise.npz"))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_determinise_isomorphic_to_reference(oracle, case):
    got = L.determinise(case["nfa"])
    assert_isomorphic(oracle, got, case["dfa"])
    assert got.start == 0 and got.hasstart and oracle.isdfa(got)
    st = L.determinise_stats()
    assert st["dfa_states"] == got.nstates and st["kernel_launches"] > 0


def test_result_executes_like_the_reference_dfa(oracle):
    """determinise -> compile -> exec on the GPU == exec of the reference's DFA (same
    language, same end-id sets per input)."""
    case = next(c for c in CASES if c["name"] == "epsunion:9pats")
    got = L.determinise(case["nfa"])
    rng = np.random.default_rng(2)
    al = np.frombuffer(b"abcdxy0123fooranch", dtype=np.uint8)
    strs = [al[rng.integers(0, al.size, int(rng.integers(0, 12)))].tobytes() for _ in range(3000)]
    strs += [b"abc", b"abd", b"axxb", b"12x", b"foobar", b"xxy", b"anch", b"abcd", b""]
    base, off = reflib.offsets_for(strs)
    want = oracle.exec_batch(case["dfa"], base, off)
    with L.Dfa(got) as dfa:
        res = dfa.exec_batch(base, off)
    assert (res["ret"] == want["ret"]).all() and (res["consumed"] == want["consumed"]).all()
    for i in np.nonzero(want["ret"] == 1)[0]:
        assert list(got.endids_of(int(res["end"][i]))) == list(case["dfa"].endids_of(int(want["end"][i])))


def test_state_limit(oracle):
    nfa = workloads.config5_nfa(30, 8, seed=3)
    n = oracle.determinise(nfa).nstates
    with pytest.raises(L.StateLimitReached):
        L.determinise(nfa, state_limit=nfa.nstates - 1)          # determinise.c:65-68
    for limit in (n - 2, n - 1, n, n + 3):
        if limit < nfa.nstates:
            continue
        want_fail = oracle.determinise(nfa, state_limit=limit) is None
        if want_fail:
            with pytest.raises(L.StateLimitReached):
                L.determinise(nfa, state_limit=limit)
        else:
            assert L.determinise(nfa, state_limit=limit).nstates == n


def test_no_start_and_not_mutating_input():
    nfa = L.FlatFsm.from_edges(3, None, [2], [(0, ord("a"), 1), (1, ord("b"), 2)])
    got = L.determinise(nfa)
    assert got.nstates == 0 and not got.hasstart


def test_config5_full_size(oracle):
    """BASELINE config 5: 100 001-state NFA -> 96.5k DFA states; isomorphic to the oracle's DFA
    (the oracle is pinned to the reference on the smaller instances of the same generator)."""
    nfa = workloads.config5_nfa(2000, 50, seed=12345)
    assert nfa.nstates == 100001
    got = L.determinise(nfa)
    st = L.determinise_stats()
    print("config5 determinise stats:", st)
    want = oracle.determinise(nfa)
    assert got.nstates == want.nstates
    assert_isomorphic(oracle, got, want)
    dense_edges = int((oracle.flatten(got) != 0xFFFFFFFF).sum())
    assert dense_edges == got.nstates * 256                         # complete: /./ self-loop at start


def test_config5_epsilon_variant_against_the_reference(oracle):
    """SURVEY.md 8d's epsilon-heavy variant: 2000 re_comp(RE_LITERAL) automata under fsm_union_array
    (201 999 states, 3998 epsilon edges).  The reference's fsm_determinise took ~150 s for it
    (tests/golden/make_golden.py main_cfg5eps); the fixture holds its state count and the sha256 of its
    canonical form, which K2's result has to reproduce."""
    g = goldenio.load_cfg5eps()
    nfa, meta = g["nfa"], g["meta"]
    assert nfa.nstates == meta["nfa_states"]
    got = L.determinise(nfa)
    st = L.determinise_stats()
    print("config5 eps-variant determinise stats:", st, "reference s:", meta["reference_determinise_s"])
    assert got.nstates == meta["dfa_states"]
    assert reflib.canonical_digest(oracle, got) == meta["dfa_canonical_sha256"]
