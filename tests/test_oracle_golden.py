"""CPU: the oracle restatement (oracle/fsm_oracle.c) against the golden fixtures that
tests/golden/make_golden.py recorded from the unmodified reference."""
import os

import numpy as np
import pytest

import goldenio
import reflib

CASES = goldenio.load_exec_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_exec.npz"))


def test_fixture_inventory():
    names = [c["name"] for c in CASES]
    assert sum(n.startswith("retest:") for n in names) >= 30
    assert any(n.startswith("cfg2:") for n in names) and any(n.startswith("utf8:") for n in names)
    assert sum(len(c["offsets"]) - 1 for c in CASES) >= 4000


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_records(oracle, case):
    f = case["fsm"]
    assert oracle.isdfa(f) == case["is_dfa"]
    if not case["is_dfa"]:
        # fsm_exec refuses with -1/EINVAL (reference src/libfsm/exec.c:106-114)
        assert (case["expect"]["ret"] == -1).all()
        assert oracle.exec(f, b"abc")[0] == -1
        return
    got = oracle.exec_batch(f, case["base"], case["offsets"], nthreads=2)
    exp, am = case["expect"], case["expect_amortised"]
    assert (got["ret"] == exp["ret"]).all()
    assert (got["consumed"] == exp["consumed"]).all()
    m = exp["ret"] == 1
    assert (got["end"][m] == exp["end"][m]).all()
    # the stop state for non-matches, from the reference's own edge_set_transition walk
    assert (got["end"] == am["end"]).all() and (got["ret"] == am["ret"]).all()
    # the reference test-suite's own '+'/'-' expectations
    t = case["tst_expect"]
    assert ((got["ret"] == 1)[t >= 0] == (t[t >= 0] == 1)).all()
    # per-call validation on == off for a DFA
    for i in range(min(3, len(case["offsets"]) - 1)):
        s = bytes(case["base"][int(case["offsets"][i]):int(case["offsets"][i + 1])])
        r = oracle.exec(f, s, validate=True)
        assert r == (int(got["ret"][i]), int(got["end"][i]), int(got["consumed"][i]))


@pytest.mark.parametrize("case", [c for c in CASES if c["is_dfa"]][::4], ids=lambda c: c["name"])
def test_oracle_flatten_matches_group_scan(oracle, case):
    f = case["fsm"]
    assert (oracle.flatten(f) == f.dense_table()).all()


def test_first_group_wins_on_ambiguous_symbol(oracle):
    # not a DFA, but edge_set_find semantics (first group in stored order) are still defined
    from libfsm_b200.desc import FlatFsm
    f = FlatFsm.from_edges(3, 0, [1], [(0, ord("a"), 2), (0, ord("a"), 1)])
    assert not oracle.isdfa(f)
    assert oracle.flatten(f)[0, ord("a")] == 1      # groups sorted by destination: 1 before 2


def test_oracle_matches_reference_on_config3_automata(oracle):
    """golden_cfg3.npz: the two BASELINE config-3 automata as the reference builds them (128-pattern
    fsm_union_repeated_pattern_group + det + min with eager outputs; rx-style anchored union with end
    ids) and the reference's answers on 2000 log lines each: records, and for the eager automaton the
    set of ids its callback received."""
    import numpy as np
    import goldenio
    g = goldenio.load_cfg3()
    e = g["eager"]
    assert g["meta"]["eager_min_states"] == e["fsm"].nstates and e["idlist"].size == 118
    got = oracle.exec_batch(e["fsm"], e["base"], e["offsets"], nthreads=4)
    assert (got == e["expect"]).all()
    ids = e["idlist"]
    for i in range(0, len(e["offsets"]) - 1, 7):
        s = e["base"][int(e["offsets"][i]):int(e["offsets"][i + 1])].tobytes()
        _, fired = oracle.exec_eager(e["fsm"], s)
        want = [int(ids[b]) for b in range(ids.size) if (int(e["masks"][i][b >> 6]) >> (b & 63)) & 1]
        assert fired == want, i
    a = g["anchored"]
    assert (oracle.exec_batch(a["fsm"], a["base"], a["offsets"], nthreads=4) == a["expect"]).all()


def test_oracle_determinise_reproduces_the_reference_on_the_epsilon_variant(oracle):
    """golden_cfg5eps.npz: 2000 re_comp literals under fsm_union_array; the reference's fsm_determinise
    result is kept as (state count, sha256 of the canonical form)."""
    g = goldenio.load_cfg5eps()
    d = oracle.determinise(g["nfa"])
    assert d.nstates == g["meta"]["dfa_states"]
    assert reflib.canonical_digest(oracle, d) == g["meta"]["dfa_canonical_sha256"]
