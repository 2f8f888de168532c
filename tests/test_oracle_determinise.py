"""CPU: the oracle's subset construction / epsilon closure against the reference's
fsm_determinise and epsilon_closure outputs recorded in tests/golden/golden_determinise.npz
(fixtures: the reference's own tests/determinise and tests/eclosure inputs, plus synthetic
NFAs).  DFAs are compared through their canonical (numbering-independent) forms."""
import os

import numpy as np
import pytest

import goldenio
import reflib
from libfsm_b200 import workloads

CASES = goldenio.load_det_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_determinise.npz"))


def assert_isomorphic(oracle, got, want):
    assert got.nstates == want.nstates, (got.nstates, want.nstates)
    a, b = reflib.canonical_form(oracle, got), reflib.canonical_form(oracle, want)
    assert a[0].shape == b[0].shape
    assert (a[0] == b[0]).all(), "transition structure differs"
    assert (a[1] == b[1]).all(), "end states differ"
    assert a[2] == b[2], "end-id sets differ"
    assert a[3] == b[3], "eager-output sets differ"


def test_fixture_inventory():
    names = [c["name"] for c in CASES]
    assert sum(n.startswith("determinise:") for n in names) == 13
    assert sum(n.startswith("eclosure:") for n in names) >= 7


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_determinise_isomorphic_to_reference(oracle, case):
    got = oracle.determinise(case["nfa"])
    assert_isomorphic(oracle, got, case["dfa"])
    assert got.start == 0 and got.hasstart


@pytest.mark.parametrize("case", [c for c in CASES if c["closure_off"] is not None], ids=lambda c: c["name"])
def test_oracle_epsilon_closure_equals_reference(oracle, case):
    off, to = oracle.epsilon_closure(case["nfa"])
    assert (off == case["closure_off"]).all() and (to == case["closure_to"]).all()


def test_state_limit_semantics(oracle):
    nfa = workloads.config5_nfa(20, 6)
    full = oracle.determinise(nfa)
    n = full.nstates
    # reference: fails when about to add a state while dfacount > limit (determinise.c:166-169),
    # and up front when the NFA itself has more states than the limit (determinise.c:65-68)
    assert oracle.determinise(nfa, state_limit=nfa.nstates - 1) is None
    big = max(n, nfa.nstates)
    assert oracle.determinise(nfa, state_limit=big).nstates == n


@pytest.mark.needs_ref
def test_state_limit_matches_reference_live(oracle, ref):
    nfa = workloads.config5_nfa(30, 8, seed=3)
    n = oracle.determinise(nfa).nstates
    for limit in (nfa.nstates, n - 2, n - 1, n, n + 5):
        if limit < nfa.nstates:
            continue
        h = ref.from_flat(nfa)
        res = ref.determinise_limit(h, limit)
        ref.free(h)
        got = oracle.determinise(nfa, state_limit=limit)
        assert (res == 1) == (got is None), (limit, res, n)


@pytest.mark.needs_ref
def test_live_differential_random_nfas(oracle, ref):
    rng = np.random.default_rng(5)
    from libfsm_b200.desc import FlatFsm
    for trial in range(25):
        n = int(rng.integers(2, 14))
        edges = [(int(rng.integers(0, n)), int(rng.integers(97, 101)), int(rng.integers(0, n))) for _ in range(int(rng.integers(1, 3 * n)))]
        eps = [(int(rng.integers(0, n)), int(rng.integers(0, n))) for _ in range(int(rng.integers(0, n)))]
        ends = sorted(set(int(x) for x in rng.integers(0, n, size=int(rng.integers(1, 4)))))
        nfa = FlatFsm.from_edges(n, 0, ends, edges, eps, endids={e: [e + 100, 7] for e in ends})
        h = ref.from_flat(nfa)
        ref.determinise(h)
        want = ref.flatten(h)
        ref.free(h)
        assert_isomorphic(oracle, oracle.determinise(nfa), want)
