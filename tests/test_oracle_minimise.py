"""CPU: the oracle's minimisation against the reference's fsm_minimise outputs recorded in
tests/golden/golden_minimise.npz (reference pipeline order: determinise, then minimise)."""
import os

import pytest

import goldenio
from test_oracle_determinise import assert_isomorphic

CASES = goldenio.load_det_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_minimise.npz"))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_minimise_isomorphic_to_reference(oracle, case):
    got = oracle.minimise(case["nfa"])          # "nfa" slot = the determinised input DFA
    want = case["dfa"]
    if want.nstates == 0:
        assert got.nstates == 0
        return
    assert_isomorphic(oracle, got, want)
    # idempotent
    assert oracle.minimise(got).nstates == got.nstates


def test_end_ids_keep_states_apart(oracle):
    case = next(c for c in CASES if c["name"] == "min:endids:union4")
    got = oracle.minimise(case["nfa"])
    ends = [tuple(got.endids_of(s)) for s in range(got.nstates) if got.is_end[s]]
    assert len(set(ends)) >= 3
