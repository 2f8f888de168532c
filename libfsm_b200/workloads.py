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


# ---- BASELINE config 3: rx(1)-style 128-pattern union over synthetic log lines ------------------

_SEV = ["ERROR", "WARN", "INFO", "DEBUG", "FATAL", "TRACE", "NOTICE", "ALERT", "CRIT", "PANIC", "AUDIT", "FAIL", "DENY", "DROP"]
_KEYS = ["user", "host", "pod", "app"]
_SVC = ["SSHD", "NGINX", "CRON", "ETCD", "REDIS", "MYSQL", "KUBE", "DOCK", "SMTP", "LDAP", "VAULT", "KAFKA"]


def cfg3_patterns():
    """The 128 PCRE patterns of BASELINE config 3 (templates such as `ERROR [0-9]{3}`,
    `user=[a-z]+`, `\\d+\\.\\d+\\.\\d+\\.\\d+`, `^\\w+ \\d{2}:\\d{2}`), UNANCHORED except two `^...` and ten
    `...$` ones, combined the reference's way for many unanchored patterns: re_comp(RE_SAVE_LINKAGE_INFO)
    -> fsm_union_repeated_pattern_group -> fsm_determinise -> fsm_minimise (tests/eager_output/utils.c),
    pattern i getting eager-output / end id i + 1.  Every template ends in a literal: the reference's
    linkage analysis (union.c:535-600) keeps a pattern that ends in a class or a repeat alive after its
    match, and the subset construction then remembers WHICH patterns have matched -- 2^n states (8
    patterns `ERROR [0-9]{3}`: 18 433 DFA states, 136 after minimise; 64 mixed ones: 2.9 M states in
    447 s) -- for the reference and, the DFA being the same, for K2 alike.  With the trailing literal
    the 128 patterns give 1590 DFA states, 1474 after minimise.
    Returns (patterns, instances): instances[i] is a byte string pattern i matches."""
    pats, inst = [], []
    forms = [(" [0-9]{3} ", " 404 "), ("\\[[0-9]+\\] ", "[17] "), (": [0-9]+ms ", ": 250ms "), ("=[0-9]{2} ", "=42 ")]
    for i in range(56):
        w = _SEV[i % len(_SEV)]
        f, ex = forms[i // len(_SEV)]
        pats.append(w + f); inst.append((w + ex).encode())
    for k in _KEYS:
        pats.append(f"{k}=[a-z]+ "); inst.append(f"{k}=alice ".encode())
    for k in ("SRC", "DST", "VIA", "GW"):
        pats.append(f"{k}=[0-9]+\\.[0-9]+\\.[0-9]+\\.[0-9]+ "); inst.append(f"{k}=10.0.12.7 ".encode())
    pats.append("^\\w+ [0-9]{2}:[0-9]{2} "); inst.append(b"Sep 12:30 ")
    pats.append("^[0-9]{4}-[0-9]{2}-[0-9]{2}T"); inst.append(b"2026-09-23T")
    for i in range(24):
        w = _SVC[i % len(_SVC)]
        f, ex = [("\\[[0-9]+\\]: ", "[812]: "), ("/[0-9]+ ", "/3 ")][i // len(_SVC)]
        pats.append(w + f); inst.append((w + ex).encode())
    for i in range(28):
        pats.append(f"CODE={100 + 7 * i} "); inst.append(f"CODE={100 + 7 * i} ".encode())
    for i in range(10):
        pats.append(f" RC={i}$"); inst.append(f" RC={i}".encode())
    assert len(pats) == len(set(pats)) == 128
    return pats, inst


def cfg3_anchored_patterns(npat: int = 128, seed: int = 7):
    """The start-anchored variant (rx(1)'s own recipe: per pattern det + min + setendid, fsm_union_array,
    fsm_determinise; end ids instead of eager outputs).  Returns (patterns, instances)."""
    rng = np.random.default_rng(seed)
    words = ["ERROR", "WARN", "INFO", "DEBUG", "FATAL", "TRACE", "kernel", "sshd", "nginx", "cron", "systemd", "postfix", "docker", "kubelet"]
    templates = [lambda w, k: f"^{w} [0-9]{{{k}}} ", lambda w, k: f"^{w}: user=[a-z]+ id=[0-9]{{{k}}}", lambda w, k: f"^{w}\\[[0-9]+\\]: ",
                 lambda w, k: f"^[0-9]{{{k}}}\\.[0-9]+\\.[0-9]+\\.[0-9]+ {w}", lambda w, k: f"^{w} (GET|POST|PUT) /[a-z/]+ ",
                 lambda w, k: f"^{w} [A-Z]{{{k}}}-[0-9]+"]
    patterns, prefixes = [], []
    while len(patterns) < npat:
        w = words[int(rng.integers(len(words)))] + str(int(rng.integers(0, 40)))
        k = int(rng.integers(1, 5)); t = int(rng.integers(len(templates)))
        p = templates[t](w, k)
        if p in patterns:
            continue
        patterns.append(p)
        ex = {0: f"{w} {'7' * k} ", 1: f"{w}: user=bob id={'4' * k}", 2: f"{w}[123]: ", 3: f"{'1' * k}.2.3.4 {w}",
              4: f"{w} GET /a/b ", 5: f"{w} {'Q' * k}-99"}[t]
        prefixes.append(ex.encode())
    return patterns, prefixes


def cfg3_lines_device(n: int, instances, seed: int = 7, lo: int = 64, hi: int = 256, device="cuda",
                      p_instance: float = 0.5, at_start: bool = False):
    """n synthetic log lines, length ~U[lo, hi], printable ASCII noise; with probability p_instance a
    line carries one pattern instance (at its start when at_start, else at a random position -- an
    instance that ends in `$`-anchored form goes to the end of the line) and, independently with the
    same probability, a second one.  Returns (base uint8 CUDA, offsets int64 CUDA [n + 1])."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lens = torch.randint(lo, hi + 1, (n,), device=device, generator=g, dtype=torch.int64)
    offsets = torch.zeros(n + 1, dtype=torch.int64, device=device)
    offsets[1:] = torch.cumsum(lens, 0)
    total = int(offsets[-1])
    base = torch.randint(0x20, 0x7F, (total,), dtype=torch.uint8, device=device, generator=g)
    pl = max(len(p) for p in instances)
    ptab = torch.zeros((len(instances), pl), dtype=torch.uint8)
    plen = torch.zeros(len(instances), dtype=torch.int64)
    is_tail = torch.zeros(len(instances), dtype=torch.bool)
    for i, p in enumerate(instances):
        ptab[i, :len(p)] = torch.tensor(list(p), dtype=torch.uint8)
        plen[i] = len(p)
        is_tail[i] = p.startswith(b" RC=")
    ptab, plen, is_tail = ptab.to(device), plen.to(device), is_tail.to(device)
    for _round in range(1 if at_start else 2):
        pid = torch.randint(0, len(instances), (n,), device=device, generator=g)
        use = torch.rand(n, device=device, generator=g) < p_instance
        room = lens - plen[pid]
        pos = (torch.rand(n, device=device, generator=g) * room.clamp(min=1).to(torch.float32)).to(torch.int64).clamp(min=0)
        if at_start:
            pos = torch.zeros_like(pos)
        pos = torch.where(is_tail[pid], room.clamp(min=0), pos)
        for k in range(pl):
            m = use & (plen[pid] > k) & (room >= 0)
            base[offsets[:-1][m] + pos[m] + k] = ptab[pid[m], k]
    return base, offsets


def cfg3_lines_host(n: int, instances, seed: int = 7, lo: int = 64, hi: int = 256, p_instance: float = 0.5,
                    at_start: bool = False):
    """Host (numpy) counterpart of cfg3_lines_device for CPU tests: same shape, its own seeded stream."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi + 1, size=n).astype(np.int64)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    base = rng.integers(0x20, 0x7F, size=int(offsets[-1]), dtype=np.uint8)
    for _round in range(1 if at_start else 2):
        for i in np.nonzero(rng.random(n) < p_instance)[0]:
            p = instances[int(rng.integers(len(instances)))]
            room = int(lens[i]) - len(p)
            if room < 0:
                continue
            pos = 0 if at_start else (room if p.startswith(b" RC=") else int(rng.integers(0, room + 1)))
            o = int(offsets[i]) + pos
            base[o:o + len(p)] = np.frombuffer(p, dtype=np.uint8)
    return base, offsets
