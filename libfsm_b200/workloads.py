"""Synthetic inputs of the BASELINE.json configurations.

Plumbing shared by tests and bench.py: seeded generators only, no automaton logic."""
from __future__ import annotations

import numpy as np

CFG2_PATTERN = r"a[ -~]{7}\z"        # -> exactly 256 states, complete
CFG1_PATTERN = r"[0-9]+\.[0-9]+"


def cfg2_host(n: int, length: int = 1024, adversarial: bool = False, seed: int = 42) -> np.ndarray:
    """[n, length] uint8: bytes uniform in [0x20, 0x7E]; adversarial: each byte 'a' w.p. 1/2."""
    rng = np.random.default_rng(seed + (1 if adversarial else 0))
    a = rng.integers(0x20, 0x7F, size=(n, length), dtype=np.uint8)
    if adversarial:
        a[rng.random((n, length), dtype=np.float32) < 0.5] = ord("a")
    return a


def cfg2_device(n: int, length: int = 1024, adversarial: bool = False, seed: int = 42, device="cuda"):
    """Same distribution generated on the device (torch), in slabs to bound temporaries."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed + (1 if adversarial else 0))
    out = torch.empty((n, length), dtype=torch.uint8, device=device)
    step = max(1, (64 << 20) // max(length, 1))
    for i in range(0, n, step):
        m = min(step, n - i)
        out[i:i + m] = torch.randint(0x20, 0x7F, (m, length), dtype=torch.uint8, device=device, generator=g)
        if adversarial:
            mask = torch.rand((m, length), device=device, generator=g) < 0.5
            out[i:i + m][mask] = ord("a")
    return out


def ragged_lines_host(n: int, lo: int = 64, hi: int = 256, seed: int = 7, alphabet: bytes | None = None):
    """n lines with length ~U[lo, hi] -> (base uint8, offsets uint64[n+1]) (config 3 shape)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi + 1, size=n).astype(np.uint64)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    if alphabet is None:
        base = rng.integers(0x20, 0x7F, size=total, dtype=np.uint8)
    else:
        al = np.frombuffer(alphabet, dtype=np.uint8)
        base = al[rng.integers(0, len(al), size=total)]
    return base, offsets


def utf8_host(nbytes: int, seed: int = 4) -> np.ndarray:
    """~nbytes of valid UTF-8: a seeded mix of 1-4 byte sequences (config 4 shape)."""
    rng = np.random.default_rng(seed)
    ncp = nbytes  # upper bound on code points
    kind = rng.choice(4, size=ncp, p=[0.7, 0.15, 0.1, 0.05])
    cp = np.empty(ncp, dtype=np.uint32)
    r = rng.random(ncp)
    cp[kind == 0] = (r[kind == 0] * 0x80).astype(np.uint32)
    cp[kind == 1] = 0x80 + (r[kind == 1] * (0x800 - 0x80)).astype(np.uint32)
    k2 = kind == 2
    v = 0x800 + (r[k2] * (0x10000 - 0x800 - 0x800)).astype(np.uint32)
    v[v >= 0xD800] += 0x800                         # skip the surrogate range
    cp[k2] = v
    cp[kind == 3] = 0x10000 + (r[kind == 3] * (0x110000 - 0x10000)).astype(np.uint32)
    nb = kind + 1
    ends = np.cumsum(nb)
    keep = int(np.searchsorted(ends, nbytes, side="right"))
    cp, nb, kind = cp[:keep], nb[:keep], kind[:keep]
    ends = ends[:keep]
    starts = ends - nb
    out = np.zeros(int(ends[-1]) if keep else 0, dtype=np.uint8)
    m = kind == 0
    out[starts[m]] = cp[m]
    m = kind == 1
    out[starts[m]] = 0xC0 | (cp[m] >> 6); out[starts[m] + 1] = 0x80 | (cp[m] & 0x3F)
    m = kind == 2
    out[starts[m]] = 0xE0 | (cp[m] >> 12); out[starts[m] + 1] = 0x80 | ((cp[m] >> 6) & 0x3F)
    out[starts[m] + 2] = 0x80 | (cp[m] & 0x3F)
    m = kind == 3
    out[starts[m]] = 0xF0 | (cp[m] >> 18); out[starts[m] + 1] = 0x80 | ((cp[m] >> 12) & 0x3F)
    out[starts[m] + 2] = 0x80 | ((cp[m] >> 6) & 0x3F); out[starts[m] + 3] = 0x80 | (cp[m] & 0x3F)
    return out


def config5_nfa(words: int = 2000, length: int = 50, seed: int = 12345):
    """BASELINE config 5's synthetic NFA: start state with a /./ self-loop plus
    `words` chains of `length` random a-z literals from start; the last state of chain w is an
    end state with end id w.  2000 x 50 -> 100 001 states (reference: 96 538 DFA states)."""
    from .desc import FlatFsm
    rng = np.random.default_rng(seed)
    letters = rng.integers(ord("a"), ord("z") + 1, size=(words, length))
    n = 1 + words * length
    nedges = 1 + words * length
    group_off = np.zeros(n + 1, dtype=np.uint64)
    sym = np.zeros((nedges, 4), dtype=np.uint64)
    to = np.zeros(nedges, dtype=np.uint32)
    # start: groups sorted by destination: self-loop (to 0) first, then chain heads ascending
    sym[0, :] = np.uint64(0xFFFFFFFFFFFFFFFF)
    to[0] = 0
    heads = 1 + np.arange(words) * length
    for w in range(words):
        c = int(letters[w, 0])
        sym[1 + w, c >> 6] = np.uint64(1) << np.uint64(c & 63)
        to[1 + w] = heads[w]
    group_off[1] = 1 + words
    g = 1 + words
    is_end = np.zeros(n, dtype=np.uint8)
    endid_off = np.zeros(n + 1, dtype=np.uint64)
    endids = []
    for w in range(words):
        for j in range(length):
            s = 1 + w * length + j
            if j + 1 < length:
                c = int(letters[w, j + 1])
                sym[g, c >> 6] = np.uint64(1) << np.uint64(c & 63)
                to[g] = s + 1
                g += 1
            else:
                is_end[s] = 1
                endids.append(w)
            group_off[s + 1] = g
            endid_off[s + 1] = len(endids)
    return FlatFsm(nstates=n, start=0, hasstart=True, is_end=is_end, group_off=group_off,
                   group_symbols=sym[:g], group_to=to[:g], eps_off=None, eps_to=None,
                   endid_off=endid_off, endids=np.array(endids, dtype=np.uint32))
