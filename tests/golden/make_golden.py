#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container,
where /root/reference exists and `make -C oracle` has produced oracle/_ref/).

Everything recorded here is an output of the reference's own code:
  * automata: re_comp -> fsm_determinise -> fsm_minimise (-> fsm_setendid, fsm_union_array)
  * expectations: fsm_exec (mode 0, the real entry point incl. its per-call validation) and
    the reference's per-byte edge_set_transition walk (mode 1, gives the stop state too)
  * vectors: the reference's own conformance vectors tests/retest/*.tst ('+' must match,
    '-' must not), plus seeded synthetic inputs.

Usage: python tests/golden/make_golden.py
"""
from __future__ import annotations

import glob
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import reflib  # noqa: E402
import goldenio  # noqa: E402
from libfsm_b200.desc import FlatFsm  # noqa: E402

REF_TESTS = "/root/reference/tests"
DIALECTS = {"pcre": reflib.RE_PCRE, "glob": reflib.RE_GLOB, "native": reflib.RE_NATIVE,
            "sql": reflib.RE_SQL, "like": reflib.RE_LIKE, "literal": reflib.RE_LITERAL}
FLAGS = {"i": 1 << 0, "t": 1 << 1, "m": 1 << 2, "r": 1 << 3, "s": 1 << 4, "z": 1 << 5, "a": 1 << 6, "x": 1 << 7}


def unescape(s: bytes) -> bytes | None:
    """parse_escapes of the reference's retest driver (src/retest/main.c:299-)."""
    out, i = bytearray(), 0
    simple = {ord('a'): 7, ord('b'): 8, ord('e'): 27, ord('f'): 12, ord('n'): 10, ord('r'): 13,
              ord('t'): 9, ord('v'): 11, ord('"'): 34, ord('\\'): 92}
    while i < len(s):
        c = s[i]
        if c != 92:
            out.append(c); i += 1; continue
        i += 1
        if i >= len(s):
            return None
        c = s[i]
        if c in simple:
            out.append(simple[c]); i += 1
        elif 48 <= c <= 55:
            v, nd = 0, 0
            while i < len(s) and nd < 3 and 48 <= s[i] <= 55:
                v = v * 8 + (s[i] - 48); nd += 1; i += 1
            out.append(v & 0xFF)
        elif c == ord('x'):
            i += 1
            curly = i < len(s) and s[i] == ord('{')
            if curly:
                i += 1
            v, nd = 0, 0
            while i < len(s) and chr(s[i]) in "0123456789abcdefABCDEF" and (curly or nd < 2):
                v = v * 16 + int(chr(s[i]), 16); nd += 1; i += 1
            if nd == 0:
                return None
            if curly:
                if i >= len(s) or s[i] != ord('}'):
                    return None
                i += 1
            out.append(v & 0xFF)
        else:
            return None
    return bytes(out)


def parse_tst(path: str):
    """Yield (dialect, flags, regexp bytes, [(should_match, input bytes), ...])."""
    dialect, flags, opts_e = "pcre", 0, False
    saved, restore = False, False
    regexp, vecs = None, []
    with open(path, "rb") as fh:
        lines = fh.read().split(b"\n")
    for raw in lines + [b""]:
        s = raw
        if len(s) == 0:
            if regexp is not None:
                yield dialect, rflags, regexp, vecs
            regexp, vecs, flags = None, [], 0
            if restore:
                opts_e = saved
            continue
        if s[:1] == b"#":
            continue
        if s[:1] == b"R" and (len(s) == 1 or s[1:2] == b" "):
            dialect = "pcre" if len(s) == 1 else s[2:].decode().strip()
            continue
        if s[:2] == b"O ":
            if s[2:3] == b"&":
                restore, saved = True, opts_e
                continue
            has_e = b"e" in s[3:]
            if s[2:3] == b"=":
                opts_e = has_e
            elif s[2:3] == b"+":
                opts_e = opts_e or has_e
            elif s[2:3] == b"-":
                opts_e = opts_e and not has_e
            continue
        if s[:2] == b"M ":
            for ch in s[2:].decode():
                if ch == "0":
                    flags = 0
                elif ch in FLAGS:
                    flags |= FLAGS[ch]
            continue
        if regexp is None:
            if s[:1] == b"~":
                s = s[1:]
            if opts_e:
                s = unescape(s)
                if s is None:
                    regexp = None
                    continue
            regexp, rflags, vecs = s, flags, []
            continue
        if s[:1] in (b"+", b"-"):
            t = unescape(s[1:])
            if t is not None:
                vecs.append((s[:1] == b"+", t))


def run_case(R, name, h, strings, tst=None, note=""):
    f = R.flatten(h)
    base, offsets = reflib.offsets_for(strings)
    exp = R.exec_batch(h, base, offsets, mode=0)
    is_dfa = not (len(strings) and exp["ret"][0] < 0)
    am = R.exec_batch(h, base, offsets, mode=1) if is_dfa else None
    if am is not None:
        m = exp["ret"] == 1
        assert (am["ret"] == exp["ret"]).all() and (am["consumed"] == exp["consumed"]).all()
        assert (am["end"][m] == exp["end"][m]).all()
    return {"name": name, "fsm": f, "base": base, "offsets": offsets, "expect": exp,
            "expect_amortised": am, "tst_expect": tst, "is_dfa": is_dfa, "note": note}


def rand_strings(rng, n, maxlen, alphabet):
    alphabet = np.frombuffer(alphabet, dtype=np.uint8)
    out = []
    for _ in range(n):
        ln = int(rng.integers(0, maxlen + 1))
        out.append(alphabet[rng.integers(0, len(alphabet), ln)].tobytes())
    return out


def cfg2_strings(rng, n, length, adversarial):
    """BASELINE config 2 input distributions."""
    a = rng.integers(0x20, 0x7F, size=(n, length), dtype=np.uint8)
    if adversarial:
        a[rng.random((n, length)) < 0.5] = ord("a")
    return [row.tobytes() for row in a]


UTF8_RE = (r"^([\x00-\x7F]|[\xC2-\xDF][\x80-\xBF]|\xE0[\xA0-\xBF][\x80-\xBF]|[\xE1-\xEC\xEE\xEF][\x80-\xBF]{2}"
           r"|\xED[\x80-\x9F][\x80-\xBF]|\xF0[\x90-\xBF][\x80-\xBF]{2}|[\xF1-\xF3][\x80-\xBF]{3}"
           r"|\xF4[\x80-\x8F][\x80-\xBF]{2})*$")


def main():
    R = reflib.Ref()
    rng = np.random.default_rng(20260922)
    cases = []

    # 1. the reference's own conformance vectors
    nvec = 0
    for path in sorted(glob.glob(os.path.join(REF_TESTS, "retest", "*.tst"))):
        for k, (dialect, flags, regexp, vecs) in enumerate(parse_tst(path)):
            if dialect not in DIALECTS or not vecs:
                continue
            try:
                h = R.compile_dfa(regexp, DIALECTS[dialect], flags, minimise=True)
            except ValueError:
                continue
            strings = [v for _, v in vecs]
            tst = np.array([1 if m else 0 for m, _ in vecs], dtype=np.int8)
            c = run_case(R, f"retest:{os.path.basename(path)}:{k}", h, strings, tst,
                         note=f"{dialect} /{regexp.decode('latin-1')}/ flags={flags}")
            # the reference's own expectation must hold for the reference's fsm_exec
            assert ((c["expect"]["ret"] == 1) == (tst == 1)).all(), c["name"]
            cases.append(c); nvec += len(vecs)
            R.free(h)
    print(f"retest: {len(cases)} regexps, {nvec} vectors")

    # 2. BASELINE config DFAs with seeded inputs (small samples of the bench workloads)
    h = R.compile_dfa(r"a[ -~]{7}\z")
    assert R.countstates(h) == 256
    cases.append(run_case(R, "cfg2:uniform", h, cfg2_strings(rng, 256, 1024, False)))
    cases.append(run_case(R, "cfg2:adversarial", h, cfg2_strings(rng, 256, 1024, True)))
    cases.append(run_case(R, "cfg2:ragged", h, rand_strings(rng, 700, 300, b"a" * 20 + bytes(range(0x20, 0x7F)) + b"\x00\n\xff")))
    cases.append(run_case(R, "cfg2:edge", h, [b"", b"a", b"a1234567", b"\x00a1234567", b"a" * 4097, b"a1234567\n"]))
    R.free(h)

    h = R.compile_dfa(r"[0-9]+\.[0-9]+")
    cases.append(run_case(R, "cfg1:digits", h, rand_strings(rng, 400, 200, b"0123456789....abc ")))
    R.free(h)

    # 3. anchored / incomplete DFAs: inputs die at assorted offsets (exec.c:133-138)
    for pat in (r"^abc[0-9]+x$", r"^(GET|POST) /[a-z]+ HTTP/1\.[01]$", r"^[a-f0-9]{32}$"):
        h = R.compile_dfa(pat)
        strs = rand_strings(rng, 300, 80, b"abcx0123456789GETPOS /HTP1.\nf")
        strs += [b"abc123x", b"abc1x", b"abcx", b"abc123", b"GET /index HTTP/1.1", b"POST /a HTTP/1.0",
                 b"GET /index HTTP/1.2", b"0123456789abcdef0123456789abcdef", b"0123456789abcdef0123456789abcdeg",
                 b"0123456789abcdef0123456789abcdef0", b""]
        cases.append(run_case(R, f"anchored:{pat}", h, strs))
        R.free(h)

    # 4. unions with end ids (tests/endids/endids2_union_many_endids.c: 6 patterns x 5 ids)
    pats = ["abc", "def", "abc.def", "abc_def", "foo", "bar"]
    hs = []
    for i, p in enumerate(pats):
        hh = R.compile_dfa(p)
        for j in range(5):
            R.setendid(hh, 1 + 5 * i + j)
        hs.append(hh)
    u = R.union_array(hs)
    R.determinise(u)
    strs = [b"abc", b"def", b"abcxdef", b"abc_def", b"foo", b"bar", b"foobar", b"xxabc_defyy", b"nothing", b"",
            b"abcdef", b"barfoo abc def"] + rand_strings(rng, 300, 40, b"abcdef_xforb ")
    cases.append(run_case(R, "endids:union6x5", u, strs))
    R.minimise(u)
    cases.append(run_case(R, "endids:union6x5:minimised", u, strs))
    R.free(u)

    # 5. a DFA with more than 256 states (16-bit table entries) built the rx(1) way, no minimise
    words = ["error", "warning", "failed", "id=[0-9]{4}", "[A-Z]{3}-[0-9]{3}", "x[0-9a-f]{3}y"]
    u = R.union_dfa(words, state_limit=20000)
    n_u = R.countstates(u)
    assert n_u > 256, n_u
    strs = rand_strings(rng, 500, 120, b"erowaningfldtmuspcby=0123456789.ABCxyf -") + \
        [b"connection failed", b"user=bob id=1234", b"warning: error", b"ABC-123", b"x0a1y", b"failed error id=0000 XYZ-999 xfffy warning"]
    cases.append(run_case(R, f"union6:{n_u}states", u, strs))
    R.free(u)

    # 6. UTF-8 validator (BASELINE config 4's automaton, as a PCRE over raw bytes)
    h = R.compile_dfa(UTF8_RE)
    good = "héllo wörld ∑ 😀 ascii".encode()
    strs = [good, good[:-1], good + b"\xff", b"\xc0\x80", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"", b"\xf0\x9f\x98\x80" * 50,
            b"\xe2\x82", b"a" * 100 + b"\x80"] + rand_strings(rng, 200, 60, bytes(range(256)))
    cases.append(run_case(R, f"utf8:{R.countstates(h)}states", h, strs))
    R.free(h)

    # 7. not a DFA: fsm_exec must refuse with -1/EINVAL (exec.c:106-114)
    h = R.re_comp(r"ab*c|abd")            # NFA with epsilons, never determinised
    cases.append(run_case(R, "notdfa:nfa", h, [b"abc", b"abd"]))
    R.free(h)

    goldenio.save_exec_cases(os.path.join(HERE, "golden_exec.npz"), cases)
    tot = sum(len(c["offsets"]) - 1 for c in cases)
    print(f"wrote golden_exec.npz: {len(cases)} cases, {tot} inputs, "
          f"{os.path.getsize(os.path.join(HERE, 'golden_exec.npz')) / 1024:.0f} KiB")


def eps_union_nfa(R, patterns, endids=True):
    """fsm_union_array of re_comp NFAs (epsilon-heavy, NOT determinised)."""
    hs = []
    for i, p in enumerate(patterns):
        h = R.re_comp(p)
        if endids:
            R.setendid(h, i)
        hs.append(h)
    return R.union_array(hs)


def det_case(R, name, h_nfa, with_closure=True, note=""):
    nfa = R.flatten(h_nfa)
    cl_off = cl_to = None
    if with_closure:
        cl_off, cl_to = R.epsilon_closure(h_nfa, nfa.nstates)
    res = R.determinise_limit(h_nfa, 20000)   # in place: h_nfa now holds the reference's DFA
    if res != 0:
        print(f"  skip {name}: fsm_determinise_with_config -> {res}")
        return None
    dfa = R.flatten(h_nfa)
    return {"name": name, "nfa": nfa, "dfa": dfa, "closure_off": cl_off, "closure_to": cl_to, "note": note}


def main_det():
    from libfsm_b200 import workloads
    R = reflib.Ref()
    cases = []
    # the reference's own fixtures: tests/determinise/in*.fsm (13) and tests/eclosure/in*.fsm (8)
    for d in ("determinise", "eclosure"):
        for path in sorted(glob.glob(os.path.join(REF_TESTS, d, "in*.fsm"))):
            h = R.parse_file(path)
            c = det_case(R, f"{d}:{os.path.basename(path)}", h)
            assert c is not None
            outp = path.replace("in", "out")
            if d == "determinise" and os.path.exists(outp):
                ho = R.parse_file(outp)
                assert R.equal(h, ho), path          # the reference's own expected output
                R.free(ho)
            if d == "eclosure":
                txt = path.replace("in", "out").replace(".fsm", ".txt")
                if os.path.exists(txt):             # exact closure sets, as `fsm -cq epsilonclosure` prints them
                    exp = {}
                    for line in open(txt):
                        k, v = line.split(":")
                        exp[int(k)] = [int(x) for x in v.split()]
                    for s2, members in exp.items():
                        got = list(c["closure_to"][int(c["closure_off"][s2]):int(c["closure_off"][s2 + 1])])
                        assert got == members, (path, s2, got, members)
            cases.append(c)
            R.free(h)
    # synthetic: BASELINE config 5's generator at small sizes, epsilon-heavy unions, regex NFAs
    for words, length in ((20, 6), (200, 10), (64, 50)):
        h = R.from_flat(workloads.config5_nfa(words, length, seed=12345))
        cases.append(det_case(R, f"cfg5:{words}x{length}", h, with_closure=False)); R.free(h); print("cfg5", words, length, flush=True)
    pats = ["^abc", "^abd", "^a.*b$", "^[0-9]+x", "^(foo|bar)+$", "^x{2,4}y", "^anch", "^a?b?c?d?$", "abd"]
    h = eps_union_nfa(R, pats)
    cases.append(det_case(R, "epsunion:9pats", h)); R.free(h)
    rng = np.random.default_rng(7)
    lits = ["^" + "".join(chr(c) for c in rng.integers(97, 103, size=int(rng.integers(2, 7)))) for _ in range(120)]
    h = eps_union_nfa(R, lits)
    cases.append(det_case(R, "epsunion:120literals", h)); R.free(h)
    for pat in (r"(a|b)*abb", r"[0-9]+\.[0-9]+", r"a[ -~]{7}\z", r"(ab|cd)*e|f+"):
        h = R.re_comp(pat)
        cases.append(det_case(R, f"re:{pat}", h)); R.free(h)
    # no start state; empty fsm
    cases = [c for c in cases if c is not None]
    goldenio.save_det_cases(os.path.join(HERE, "golden_determinise.npz"), cases)
    print(f"wrote golden_determinise.npz: {len(cases)} cases, "
          f"{os.path.getsize(os.path.join(HERE, 'golden_determinise.npz')) / 1024:.0f} KiB, "
          f"largest DFA {max(c['dfa'].nstates for c in cases)} states")


def main_min():
    """golden_minimise.npz: (reference-determinised DFA, the reference's fsm_minimise of it), in
    the reference's own pipeline order (determinise then minimise on the same struct fsm, as
    re(1)/lx(1)/retest do: src/re/main.c:872-881)."""
    R = reflib.Ref()
    cases = []
    det = goldenio.load_det_cases(os.path.join(HERE, "golden_determinise.npz"))
    for c in det:
        h = R.from_flat(c["nfa"])
        R.determinise(h)
        dfa_in = R.flatten(h)
        R.minimise(h)
        cases.append({"name": "min:" + c["name"], "nfa": dfa_in, "dfa": R.flatten(h)})
        R.free(h)
    for pat in (r"a[ -~]{7}\z", r"[0-9]+\.[0-9]+", r"^(GET|POST) /[a-z]+ HTTP/1\.[01]$", r"(a|b)*abb(a|b)*", r"x{3,5}y|x{4}z", "(foo|bar)+baz|qux"):
        h = R.re_comp(pat.replace("\\\\", "\\"))
        R.determinise(h)
        dfa_in = R.flatten(h)
        R.minimise(h)
        cases.append({"name": "min:re:" + pat, "nfa": dfa_in, "dfa": R.flatten(h)})
        R.free(h)
    # unions with end ids: states that differ only in their end-id sets must stay apart
    hs = []
    for i, p in enumerate(["abc", "abd", "ab.", "xyz"]):
        hh = R.compile_dfa(p); R.setendid(hh, 10 + i); hs.append(hh)
    u = R.union_array(hs)
    R.determinise(u)
    dfa_in = R.flatten(u)
    R.minimise(u)
    cases.append({"name": "min:endids:union4", "nfa": dfa_in, "dfa": R.flatten(u)})
    R.free(u)
    goldenio.save_det_cases(os.path.join(HERE, "golden_minimise.npz"), cases)
    print(f"wrote golden_minimise.npz: {len(cases)} cases, {os.path.getsize(os.path.join(HERE, 'golden_minimise.npz')) / 1024:.0f} KiB, "
          f"largest input {max(c['nfa'].nstates for c in cases)} -> {max(c['dfa'].nstates for c in cases)} states")


def main_fixtures():
    """golden_re_fixtures.npz: the reference's regex golden files (tests/<dialect>/inN.re ->
    outN.fsm, compared by the reference with `fsm -t equal`), as (dialect, re(1) arguments, regex
    bytes, expected automaton).  Kept only if the reference's own re(1) reproduces outN.fsm here."""
    import subprocess
    import tempfile
    R = reflib.Ref()
    re_ref = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "re_ref")
    dirs = [("pcre", "pcre", []), ("pcre-anchor", "pcre", []), ("pcre-repeat", "pcre", []), ("pcre-flags", "pcre", ["-b"]),
            ("native", "native", []), ("glob", "glob", []), ("like", "like", []), ("literal", "literal", []), ("sql", "sql", [])]
    out, meta, kept, skipped = {}, [], 0, 0
    for d, dialect, extra in dirs:
        for path in sorted(glob.glob(os.path.join(REF_TESTS, d, "in*.re"))):
            n = os.path.basename(path)[2:-3]
            outp = os.path.join(REF_TESTS, d, f"out{n}.fsm")
            if not os.path.exists(outp):
                continue
            args = list(extra)
            modep = os.path.join(REF_TESTS, d, f"mode{n}")
            if os.path.exists(modep):
                args = ["-F", open(modep).read().strip()] + args
            regex = open(path, "rb").read()
            with tempfile.TemporaryDirectory() as td:
                rf = os.path.join(td, "in.re"); open(rf, "wb").write(regex)
                p = subprocess.run([re_ref] + args + ["-r", dialect, "-py", rf], capture_output=True, timeout=60)
                if p.returncode != 0:
                    skipped += 1; continue
                gf = os.path.join(td, "got.fsm"); open(gf, "wb").write(p.stdout)
                hg, he = R.parse_file(gf), R.parse_file(outp)
                same = R.equal(hg, he)
                exp = R.flatten(he)
                R.free(hg); R.free(he)
            if not same:
                skipped += 1; continue
            pre = f"f{kept}_"
            goldenio.pack_fsm(pre, exp, out)
            out[pre + "regex"] = np.frombuffer(regex, dtype=np.uint8)
            meta.append({"name": f"{d}:in{n}", "dialect": dialect, "args": args})
            kept += 1
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "golden_re_fixtures.npz"), **out)
    print(f"wrote golden_re_fixtures.npz: {kept} fixtures ({skipped} skipped), "
          f"{os.path.getsize(os.path.join(HERE, 'golden_re_fixtures.npz')) / 1024:.0f} KiB")


def main_eager():
    """golden_eager.npz: eager outputs (include/fsm/fsm.h:273-336) through the reference's own
    pipeline: NFA with eager ids -> fsm_determinise -> fsm_minimise -> fsm_exec with the callback.
    Random NFAs (the generator of tests/test_oracle_eager.py) and unions built by the reference's
    fsm_union_repeated_pattern_group, the way tests/eager_output/utils.c:run_test does."""
    sys.path.insert(0, os.path.dirname(HERE))
    from test_oracle_eager import diamond, random_nfa
    R = reflib.Ref()
    cases = []

    def record(name, h, nfa, inputs):
        R.determinise(h)
        dfa = R.flatten(h)
        R.minimise(h)
        mn = R.flatten(h)
        fired, rets = [], []
        for s in inputs:
            (ret, _end, _consumed), ids = R.exec_eager(h, s)
            fired.append(ids); rets.append(int(ret))
        cases.append({"name": name, "nfa": nfa, "dfa": dfa, "min": mn if mn.nstates else None,
                      "inputs": inputs, "fired": fired, "rets": rets})
        R.free(h)

    al = np.frombuffer(b"abcdx", dtype=np.uint8)
    for seed in range(16):
        rng = np.random.default_rng(9000 + seed)
        nfa = random_nfa(rng, int(rng.integers(4, 18)))
        inputs = [al[rng.integers(0, al.size, int(rng.integers(0, 10)))].tobytes() for _ in range(24)] + [b""]
        record(f"random:{seed}", R.from_flat(nfa), nfa, inputs)
    for k, eager in enumerate([{1: [7]}, {2: [7]}, {1: [7], 2: [8]}]):
        f = diamond(eager)
        record(f"blindspot:{k}", R.from_flat(f), f, [b"ac", b"bc", b"a", b"b", b"", b"ab"])
    RE_SAVE_LINKAGE_INFO = 1 << 9
    for name, pats, inputs in (
        ("group:abc,b+,xyz", ["abc", "b+", "xyz"], [b"abc", b"zabcz", b"bbb", b"xyzabc", b"", b"q"]),
        ("group:anchors", ["^ab", "cd$", "e"], [b"ab", b"xab", b"cd", b"cdx", b"abecd", b"e"]),
        ("group:overlap", ["a+b", "ab+", "b", "[ab]{3}"], [b"ab", b"aab", b"abb", b"bbb", b"aaa", b"ba"]),
    ):
        hs = [R.re_comp(p, flags=RE_SAVE_LINKAGE_INFO) for p in pats]
        u = R.union_repeated_pattern_group(hs, 1)
        record(name, u, R.flatten(u), inputs)
    goldenio.save_eager_cases(os.path.join(HERE, "golden_eager.npz"), cases)
    print(f"wrote golden_eager.npz: {len(cases)} cases, {os.path.getsize(os.path.join(HERE, 'golden_eager.npz')) / 1024:.0f} KiB")


def main_cfg3():
    """golden_cfg3.npz: BASELINE config 3's two automata as the REFERENCE builds them, plus a sample of
    lines with the reference's own answers.
      eager    libfsm_b200.workloads.cfg3_patterns(): 128 patterns, unanchored but for 2 `^` and 10 `$`
               ones: re_comp(RE_SAVE_LINKAGE_INFO) -> fsm_union_repeated_pattern_group(id_base 1) ->
               fsm_determinise -> fsm_minimise (tests/eager_output/utils.c:67-131); per line the
               reference's fsm_exec record and the set of eager-output ids its callback received
      anchored cfg3_anchored_patterns(): rx(1)'s recipe (src/rx/main.c:487-566,1353,1371: per pattern
               det + min + setendid(i), fsm_union_array, fsm_determinise, no final minimise)"""
    from libfsm_b200 import workloads
    R = reflib.Ref()
    out = {}
    RE_SAVE_LINKAGE_INFO = 1 << 9
    pats, inst = workloads.cfg3_patterns()
    hs = [R.re_comp(p, reflib.RE_PCRE, RE_SAVE_LINKAGE_INFO) for p in pats]
    u = R.union_repeated_pattern_group(hs, 1)
    R.determinise(u)
    ndet = R.countstates(u)
    R.minimise(u)
    f = R.flatten(u)
    goldenio.pack_fsm("eager_", f, out)
    base, off = workloads.cfg3_lines_host(2000, inst, seed=11)
    ids = np.unique(f.eager_ids)
    rec, masks = R.exec_eager_batch(u, base, off, ids, mode=0, nthreads=8)
    out.update(eager_base=base, eager_offsets=off, eager_expect=rec.view(np.uint8), eager_masks=masks, eager_idlist=ids)
    R.free(u)
    apats, ainst = workloads.cfg3_anchored_patterns()
    h = R.union_dfa(apats, state_limit=500000)
    fa = R.flatten(h)
    goldenio.pack_fsm("anch_", fa, out)
    base, off = workloads.cfg3_lines_host(2000, ainst, seed=12, at_start=True)
    rec = R.exec_batch(h, base, off, mode=1, nthreads=8)
    asis = R.exec_batch(h, base, off, mode=0, nthreads=8)
    assert (asis["ret"] == rec["ret"]).all() and (asis["consumed"] == rec["consumed"]).all()
    out.update(anch_base=base, anch_offsets=off, anch_expect=rec.view(np.uint8))
    R.free(h)
    out["meta"] = np.frombuffer(json.dumps({"eager_det_states": ndet, "eager_min_states": f.nstates,
                                            "anch_states": fa.nstates}).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "golden_cfg3.npz")
    np.savez_compressed(path, **out)
    print(f"wrote golden_cfg3.npz: eager {ndet} -> {f.nstates} states ({ids.size} eager ids), anchored {fa.nstates} states, "
          f"{os.path.getsize(path) / 1024:.0f} KiB")


def main_cfg4():
    """golden_cfg4.npz: BASELINE config 4's validator built the way SURVEY.md 8d says: examples/utf8dfa
    fed `0..10FFFF` (restated over the same API calls in oracle/ref_harness.c: refh_utf8dfa -- the example
    can only print dot / api / c) gives the 9-state DFA of ONE code point; epsilon edges from its end
    states back to the start state + start state accepting (refh_star), fsm_determinise, fsm_minimise give
    the 8-state stream validator.  fsm_equal with the PCRE-built validator of golden_exec.npz ("utf8:")
    is recorded in the meta (and asserted here)."""
    R = reflib.Ref()
    h = R.utf8dfa(0, 0x10FFFF)
    one_states = R.countstates(h)
    R.star(h); R.determinise(h); R.minimise(h)
    f = R.flatten(h)
    cases = {c["name"]: c for c in goldenio.load_exec_cases(os.path.join(HERE, "golden_exec.npz"))}
    pname = next(n for n in cases if n.startswith("utf8:"))
    p = R.from_flat(cases[pname]["fsm"])
    eq = R.equal(h, p)
    assert eq and one_states == 9 and f.nstates == 8
    out = {}
    goldenio.pack_fsm("utf8dfa_star_", f, out)
    out["meta"] = np.frombuffer(json.dumps({"one_codepoint_states": one_states, "validator_states": f.nstates,
                                            "fsm_equal_with": pname, "fsm_equal": bool(eq)}).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "golden_cfg4.npz"), **out)
    print(f"wrote golden_cfg4.npz: utf8dfa 0..10FFFF -> {one_states} states, starred + det + min -> {f.nstates} states, fsm_equal({pname}) = {eq}")
    R.free(h); R.free(p)


canonical_digest = reflib.canonical_digest


def main_cfg5eps():
    """golden_cfg5eps.npz: BASELINE config 5's epsilon-heavy variant: 2000 random 50-letter
    literals, each compiled by re_comp(RE_LITERAL) and given end id = its index, combined with
    fsm_union_array (a new start state with an epsilon edge to every sub-automaton; 3998 epsilon edges,
    201 999 states).  The reference's fsm_determinise needs ~150 s for it on this box; its result is kept as
    (state count, sha256 of the canonical form) -- the DFA itself would be 90 MB."""
    R = reflib.Ref()
    O = reflib.Oracle()
    rng = np.random.default_rng(12345)
    words = ["".join(chr(c) for c in rng.integers(97, 123, size=50)) for _ in range(2000)]
    hs = [R.re_comp(w, reflib.RE_LITERAL, 0) for w in words]
    for i, h in enumerate(hs):
        R.setendid(h, i)
    u = R.union_array(hs)
    nfa = R.flatten(u)
    import time
    t0 = time.time(); R.determinise(u); dt = time.time() - t0
    dfa = R.flatten(u)
    out = {}
    goldenio.pack_fsm("nfa_", nfa, out)
    out["meta"] = np.frombuffer(json.dumps({"nfa_states": nfa.nstates, "eps_edges": int(nfa.eps_off[-1]), "dfa_states": dfa.nstates,
                                            "dfa_canonical_sha256": canonical_digest(O, dfa), "reference_determinise_s": dt}).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "golden_cfg5eps.npz")
    np.savez_compressed(path, **out)
    print(f"wrote golden_cfg5eps.npz: NFA {nfa.nstates} states / {int(nfa.eps_off[-1])} eps edges -> DFA {dfa.nstates} states in {dt:.1f} s (reference), "
          f"{os.path.getsize(path) / 1024:.0f} KiB")
    R.free(u)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "cfg5eps":       # ~3 min: not part of the default regeneration
        main_cfg5eps()
    if len(sys.argv) > 1 and sys.argv[1] == "cfg4":          # ~40 s: not part of the default regeneration
        main_cfg4()
    if len(sys.argv) < 2 or sys.argv[1] == "cfg3":
        main_cfg3()
    if len(sys.argv) < 2 or sys.argv[1] == "eager":
        main_eager()
    if len(sys.argv) < 2 or sys.argv[1] == "fixtures":
        main_fixtures()
    if len(sys.argv) < 2 or sys.argv[1] == "min":
        main_min()
    if len(sys.argv) < 2 or sys.argv[1] == "exec":
        main()
    if len(sys.argv) < 2 or sys.argv[1] == "det":
        main_det()
