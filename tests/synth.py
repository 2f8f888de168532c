"""Synthetic DFAs with a prescribed byte-class structure (test helper: seeded, no automaton logic
beyond laying out random transitions)."""
import numpy as np

from libfsm_b200.desc import FlatFsm


def dfa_from_classes(class_of, nstates, seed, missing=0.0, ends=0.3):
    """A DFA over `nstates` states whose transition depends on the byte only through class_of[byte]
    (class_of: 256 small ints).  `missing`: probability that a (state, class) pair has no edge."""
    rng = np.random.default_rng(seed)
    class_of = np.asarray(class_of, dtype=np.int64)
    ncls = int(class_of.max()) + 1
    # every class gets a distinct column in at least one row, so the engine finds exactly these classes
    nxt = rng.integers(0, nstates, size=(nstates, ncls))
    nxt[0] = (np.arange(ncls) + 1) % nstates if nstates > ncls else nxt[0]
    gone = rng.random((nstates, ncls)) < missing
    gone[0] = False
    edges = []
    for s in range(nstates):
        for c in range(ncls):
            if gone[s, c]:
                continue
            syms = [int(b) for b in np.nonzero(class_of == c)[0]]
            if syms:
                edges.append((s, syms, int(nxt[s, c])))
    end_states = [s for s in range(nstates) if rng.random() < ends] or [nstates - 1]
    return FlatFsm.from_edges(nstates, 0, end_states, edges)


def classes_from_ranges(r0, r1=None, extra=None):
    """class_of[256] for the partition induced by up to two byte ranges (lo, hi inclusive)."""
    c = np.zeros(256, dtype=np.int64)
    if r0 is not None:
        c[r0[0]:r0[1] + 1] |= 1
    if r1 is not None:
        c[r1[0]:r1[1] + 1] |= 2
    if extra is not None:
        c[extra[0]:extra[1] + 1] |= 4
    _, inv = np.unique(c, return_inverse=True)
    return inv
