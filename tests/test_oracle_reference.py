"""CPU: live differential of the oracle restatement against the compiled reference
(oracle/_ref/libref_harness.so).  Skipped where the reference could not be built."""
import numpy as np
import pytest

import reflib

PATTERNS = [r"a[ -~]{7}\z", r"[0-9]+\.[0-9]+", r"^abc[0-9]+x$", r"(foo|bar)+baz", r"^$", r"a*b*c*",
            r"^[a-z]+@[a-z]+\.(com|org)$", r"x{3,5}y", r"(?i)hello", r"[^a]b"]


@pytest.mark.parametrize("pattern", PATTERNS)
def test_exec_random_inputs(oracle, ref, pattern):
    h = ref.compile_dfa(pattern)
    f = ref.flatten(h)
    rng = np.random.default_rng(abs(hash(pattern)) % (2 ** 32))
    alpha = np.frombuffer(b"abcxyz0123456789.@fobarhelHELO \n\x00", dtype=np.uint8)
    strs = [alpha[rng.integers(0, len(alpha), int(rng.integers(0, 64)))].tobytes() for _ in range(2000)]
    base, off = reflib.offsets_for(strs)
    got = oracle.exec_batch(f, base, off)
    exp0 = ref.exec_batch(h, base, off, mode=0)
    exp1 = ref.exec_batch(h, base, off, mode=1)
    assert (got["ret"] == exp0["ret"]).all() and (got["consumed"] == exp0["consumed"]).all()
    assert (got["end"][exp0["ret"] == 1] == exp0["end"][exp0["ret"] == 1]).all()
    assert (got == exp1).all()
    ref.free(h)


def test_round_trip_desc(oracle, ref):
    """flatten -> refh_from_desc -> flatten is the identity, and fsm_equal agrees."""
    h = ref.compile_dfa(r"(ab|cd)*e")
    f = ref.flatten(h)
    h2 = ref.from_flat(f)
    f2 = ref.flatten(h2)
    assert ref.equal(h, h2)
    assert (f.dense_table() == f2.dense_table()).all() and (f.is_end == f2.is_end).all()
    ref.free(h); ref.free(h2)


def test_nfa_rejected_like_reference(oracle, ref):
    h = ref.re_comp(r"ab*c|abd")
    f = ref.flatten(h)
    assert ref.exec(h, b"abc")[0] == -1
    assert oracle.exec(f, b"abc")[0] == -1
    ref.free(h)


def test_endids_match_reference(oracle, ref):
    hs = []
    for i, p in enumerate(["abc", "def", "abc.def"]):
        hh = ref.compile_dfa(p)
        ref.setendid(hh, 10 + i)
        hs.append(hh)
    u = ref.union_array(hs)
    ref.determinise(u)
    f = ref.flatten(u)
    for s in (b"abc", b"abcxdef", b"def", b"zzz"):
        rc, end, _ = ref.exec(u, s)
        ret, oend, _ = oracle.exec(f, s)
        assert rc == ret
        if rc == 1:
            assert end == oend
            assert list(f.endids_of(end)) == ref.endids(u, end)
    ref.free(u)
