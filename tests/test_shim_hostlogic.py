"""CPU: the libfsm-side shim's HOST logic, without a GPU.

build/shim_cpu/ holds the UNCHANGED shim source (libfsm_b200/shim/fsm_b200_shim.c) linked with
the reference's objects and, in place of libfsm_b200.so, oracle/stub_engine.c -- the plain-C
oracle answering the engine's symbols (test infrastructure; `make -C libfsm_b200/shim STUB=1`).
Everything the shim does on the host is therefore exercised here by the reference's own CLIs and
C unit tests: struct fsm -> flat description, the compiled-table cache (fingerprints, LRU,
pinning under threads), getc draining and cursor restoration, the struct fsm rebuild after
determinise / minimise, errno conventions.  The same programs run against the CUDA engine in
tests/test_gpu_shim.py.
"""
import os
import subprocess

import numpy as np
import pytest

import goldenio
import reflib
from test_gpu_shim import FIXTURES_NPZ, RE_REF, FSM_REF, to_fsm5

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPU_DIR = os.path.join(ROOT, "build", "shim_cpu")
RE_CPU = os.path.join(CPU_DIR, "re_b200")
FSM_CPU = os.path.join(CPU_DIR, "fsm_b200")


def _build():
    if os.path.exists("/root/reference/src/libfsm/exec.c") and os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "obj")):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "libfsm_b200", "shim"), "STUB=1", "-j8"], check=True,
                       stdout=subprocess.DEVNULL)


_build()
pytestmark = pytest.mark.skipif(not (os.path.exists(RE_CPU) and os.path.exists(RE_REF)),
                                reason="build/shim_cpu not built (needs the reference tree at build time)")


def run(binary, args, **kw):
    p = subprocess.run([binary] + args, capture_output=True, timeout=120, **kw)
    return p.returncode, p.stdout, p.stderr


def test_selftest_program():
    p = subprocess.run([os.path.join(CPU_DIR, "shim_selftest")], capture_output=True, timeout=120)
    assert p.returncode == 0, (p.stdout.decode(), p.stderr.decode())
    assert b"shim selftest ok" in p.stdout


@pytest.mark.parametrize("threads,rounds,stall_us", [(4, 40, 0), (32, 5, 200)])
def test_table_cache_is_safe_under_threads(threads, rounds, stall_us):
    """More threads than cache slots + a stall inside every engine call: before the cache pinned
    entries (refs), an entry could be evicted and freed while another thread was executing on it
    (the stub engine detects the use after free and the verdicts differ)."""
    env = dict(os.environ)
    if stall_us:
        env["FSM_B200_STUB_SLOW"] = str(stall_us)
    p = subprocess.run([os.path.join(CPU_DIR, "shim_threads"), str(threads), str(rounds)], capture_output=True,
                       timeout=300, env=env)
    assert p.returncode == 0, (p.stdout.decode(), p.stderr.decode())


REFTESTS_DIR = os.path.join(CPU_DIR, "reftests")
REFTESTS = sorted(os.listdir(REFTESTS_DIR)) if os.path.isdir(REFTESTS_DIR) else []


@pytest.mark.parametrize("name", REFTESTS)
def test_reference_own_c_unit_tests(name):
    """tests/endids/*.c and tests/re_strings/*.c of the reference, compiled unmodified."""
    p = subprocess.run([os.path.join(REFTESTS_DIR, name)], capture_output=True, timeout=300)
    assert p.returncode == 0, (name, p.stdout.decode()[-2000:], p.stderr.decode()[-2000:])


def test_config1_files(tmp_path):
    rng = np.random.default_rng(1)
    match = rng.integers(ord("0"), ord("9") + 1, size=1 << 20, dtype=np.uint8)
    match[rng.random(match.size) < 1 / 64] = ord(".")
    nomatch = rng.integers(ord("a"), ord("z") + 1, size=1 << 20, dtype=np.uint8)
    fm, fn = tmp_path / "match.txt", tmp_path / "nomatch.txt"
    fm.write_bytes(match.tobytes()); fn.write_bytes(nomatch.tobytes())
    for files in ([str(fm)], [str(fn)], [str(fm), str(fn)], [str(fn), str(fm)]):
        args = ["-r", "pcre", "-x", r"[0-9]+\.[0-9]+"] + files
        got, want = run(RE_CPU, args), run(RE_REF, args)
        assert got[0] == want[0] and got[1] == want[1], (files, got, want)


@pytest.mark.parametrize("args", [
    ["-r", "pcre", r"a[ -~]{7}\z", "xxabcdefgh"],
    ["-r", "pcre", r"a[ -~]{7}\z", "xxabcdefg"],
    ["-r", "pcre", r"^abc[0-9]+x$", "abc123x", "abc12", "zzz"],
    ["-r", "native", "ab*c", "abbbc"],
    ["-r", "glob", "*.txt", "notes.txt"],
    ["-r", "literal", "hello", "hello"],
    ["-r", "pcre", "-z", "abc", "def", "xyz"],
])
def test_argv_strings_same_exit_status_and_output(args):
    got, want = run(RE_CPU, args), run(RE_REF, args)
    assert got[0] == want[0] and got[1] == want[1], (args, got, want)


def test_stdin_stream_cursor(tmp_path):
    """fsm_fgetc path: `re -x PATTERN FILE` reads FILE through fsm_fgetc; a miss must leave the
    exit status and output the reference's."""
    f = tmp_path / "in.txt"
    f.write_bytes(b"abc123\nzzz\n")
    for pat in (r"abc[0-9]+", r"^zzz", r"nomatch"):
        args = ["-r", "pcre", "-x", pat, str(f)]
        got, want = run(RE_CPU, args), run(RE_REF, args)
        assert got[0] == want[0] and got[1] == want[1], (pat, got, want)


def test_fsm_cli_determinise_and_minimise(tmp_path):
    """`fsm -pd` / `fsm -pm` of the relinked fsm(1) against the reference's, compared the way
    the reference's own tests do (tests/determinise/Makefile:11-20: fsm -t equal) + state counts."""
    cases = goldenio.load_det_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_determinise.npz"))
    ran = 0
    for c in cases:
        txt = to_fsm5(c["nfa"])
        if txt is None or c["dfa"].nstates > 400:
            continue
        inp = tmp_path / "in.fsm"
        inp.write_text(txt)
        for flag in ("-pd", "-pm"):
            outs = {}
            for tag, binary in (("cpu", FSM_CPU), ("ref", FSM_REF)):
                p = subprocess.run([binary, flag], stdin=open(inp), capture_output=True, timeout=120)
                assert p.returncode == 0, (c["name"], flag, tag, p.stderr)
                outs[tag] = tmp_path / f"out_{tag}.fsm"
                outs[tag].write_bytes(p.stdout)
            eq = subprocess.run([FSM_REF, "-t", "equal", str(outs["cpu"]), str(outs["ref"])], capture_output=True, timeout=120)
            assert eq.returncode == 0, (c["name"], flag, eq.stdout, eq.stderr)
            cnt = [subprocess.run([FSM_REF, "-q", "count"], stdin=open(outs[t]), capture_output=True).stdout for t in ("cpu", "ref")]
            assert cnt[0] == cnt[1], (c["name"], flag)
        ran += 1
    assert ran >= 10


def test_reference_regex_golden_files(tmp_path):
    """All of the reference's regex golden-file fixtures (tests/pcre*, native, glob, like, literal,
    sql: `re -r D -py inN.re` vs outN.fsm by fsm_equal) through the relinked re(1)."""
    if not os.path.exists(FIXTURES_NPZ) or not reflib.have_ref():
        pytest.skip("fixtures or compiled reference missing")
    ref = reflib.Ref()
    fixtures = goldenio.load_re_fixtures(FIXTURES_NPZ)
    assert len(fixtures) >= 200
    bad = []
    for fx in fixtures:
        rf = tmp_path / "in.re"
        rf.write_bytes(fx["regex"])
        p = subprocess.run([RE_CPU] + fx["args"] + ["-r", fx["dialect"], "-py", str(rf)], capture_output=True, timeout=120)
        if p.returncode != 0:
            bad.append((fx["name"], "exit", p.returncode, p.stderr[-200:]))
            continue
        gf = tmp_path / "got.fsm"
        gf.write_bytes(p.stdout)
        hg = ref.parse_file(str(gf))
        he = ref.from_flat(fx["fsm"])
        if not ref.equal(hg, he):
            bad.append((fx["name"], "language differs"))
        ref.free(hg); ref.free(he)
    assert not bad, bad[:10]


def test_fsm_cli_text_is_the_reference_text_with_reference_numbering(tmp_path):
    """FSM_B200_DET_NUMBERING=reference: the engine returns the DFA in the reference's own state
    numbering (here: the product's refnum.h functions run on the CPU, oracle/refnum_host.cpp), the
    shim rebuilds the struct fsm, and `fsm -pd` prints BYTE FOR BYTE what the reference's fsm(1)
    prints -- no `fsm -t equal`, no canonicalisation."""
    cases = goldenio.load_det_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_determinise.npz"))
    env = dict(os.environ, FSM_B200_DET_NUMBERING="reference")
    ran = differs_without = 0
    for c in cases:
        txt = to_fsm5(c["nfa"])
        if txt is None or c["dfa"].nstates > 3500:
            continue
        inp = tmp_path / "in.fsm"
        inp.write_text(txt)
        got = subprocess.run([FSM_CPU, "-pd"], stdin=open(inp), capture_output=True, timeout=120, env=env)
        bfs = subprocess.run([FSM_CPU, "-pd"], stdin=open(inp), capture_output=True, timeout=120)
        want = subprocess.run([FSM_REF, "-pd"], stdin=open(inp), capture_output=True, timeout=120)
        assert got.returncode == 0 and want.returncode == 0, (c["name"], got.stderr)
        assert got.stdout == want.stdout, c["name"]
        differs_without += bfs.stdout != want.stdout
        ran += 1
    assert ran >= 15
    assert differs_without > 0      # the flag matters: BFS numbering prints different text somewhere


LX_CPU = os.path.join(CPU_DIR, "lx_b200")
LX_REF = os.path.join(ROOT, "oracle", "_ref", "lx_ref")


@pytest.mark.skipif(not (os.path.exists(LX_CPU) and os.path.exists(LX_REF) and os.path.isdir("/root/reference/src")),
                    reason="relinked lx(1) / reference tree not available")
def test_lx_relinked_generates_equivalent_lexers(tmp_path):
    """lx(1) relinked, unchanged, against the shim (run with -C 1, see lxcheck.token_stream).
    State numbers in the generated code differ
    (minimal DFAs are unique up to numbering), so the check is behavioural: the C lexer it generates
    from the reference's own .lx specifications must tokenise like the one the reference's lx
    generates -- same tokens, same spellings, same positions."""
    from lxcheck import SPECS, token_stream
    for spec, text in SPECS:
        path = os.path.join("/root/reference", spec)
        if not os.path.exists(path):
            continue
        got = token_stream(LX_CPU, path, text, tmp_path / "cpu")
        want = token_stream(LX_REF, path, text, tmp_path / "ref")
        assert got == want, spec
        assert want.count(b"\n") >= 3


@pytest.mark.skipif(not (os.path.exists(LX_CPU) and os.path.exists(LX_REF)), reason="relinked lx(1) not built")
def test_lx_relinked_on_the_repository_sample_spec(tmp_path):
    """Same check on tests/data/sample.lx (written for this repository, so it also runs where the
    reference tree is absent)."""
    from lxcheck import SAMPLE_SPEC, SAMPLE_TEXT, token_stream
    got = token_stream(LX_CPU, SAMPLE_SPEC, SAMPLE_TEXT, tmp_path / "cpu")
    want = token_stream(LX_REF, SAMPLE_SPEC, SAMPLE_TEXT, tmp_path / "ref")
    assert got == want and want.count(b"\n") == 37


def test_shim_is_thread_sanitizer_clean(tmp_path):
    """The shim + stub engine + thread stress rebuilt with -fsanitize=thread (reference objects as
    they are): no data race may be reported in fsm_b200_shim.c.  Skips where libtsan is missing."""
    ref_root = "/root/reference"
    obj_dir = os.path.join(ROOT, "oracle", "_ref", "obj")
    if not (os.path.isdir(os.path.join(ref_root, "src")) and os.path.isdir(obj_dir)):
        pytest.skip("needs the reference tree and its compiled objects")
    objs = []
    for d, _, files in os.walk(obj_dir):
        for f in sorted(files):
            if f.endswith(".o") and not (d.endswith("libfsm") and f in ("exec.o", "determinise.o", "minimise.o")):
                objs.append(os.path.join(d, f))
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + ref_root + "/include", "-I" + ref_root + "/src", "-I" + ref_root + "/src/libfsm"]
    so = str(tmp_path / "libfsm_shim.so")
    p = subprocess.run(["gcc", "-std=gnu99", "-O1", "-g", "-fsanitize=thread", "-fPIC", "-w", "-shared", "-pthread", "-o", so,
                        os.path.join(ROOT, "libfsm_b200", "shim", "fsm_b200_shim.c"), os.path.join(ROOT, "oracle", "stub_engine.c"),
                        os.path.join(ROOT, "oracle", "fsm_oracle.c")] + inc + objs +
                       ["-L" + os.path.join(ROOT, "oracle", "_ref"), "-lrefnum_host", "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_ref"),
                        "-Wl,-Bsymbolic"], capture_output=True, timeout=300)
    if p.returncode != 0:
        pytest.skip("ThreadSanitizer build not available here: " + p.stderr.decode()[-200:])
    exe = str(tmp_path / "shim_threads")
    subprocess.run(["gcc", "-std=c99", "-O1", "-g", "-fsanitize=thread", "-pthread", "-w", "-D_XOPEN_SOURCE=600"] + inc +
                   ["-I" + os.path.join(ROOT, "libfsm_b200", "shim"), "-o", exe, os.path.join(ROOT, "libfsm_b200", "shim", "shim_threads.c"),
                    "-L" + str(tmp_path), "-lfsm_shim", "-Wl,-rpath," + str(tmp_path)], check=True, timeout=300)
    r = subprocess.run([exe, "24", "3"], capture_output=True, timeout=600, env=dict(os.environ, FSM_B200_STUB_SLOW="100"))
    assert r.returncode == 0, (r.stdout.decode()[-500:], r.stderr.decode()[-2000:])
    assert b"ThreadSanitizer" not in r.stderr, r.stderr.decode()[:3000]


RX_CPU = os.path.join(CPU_DIR, "rx_b200")
RX_REF = os.path.join(ROOT, "oracle", "_ref", "rx_ref")


@pytest.mark.skipif(not (os.path.exists(RX_CPU) and os.path.exists(RX_REF)), reason="relinked rx(1) not built")
def test_rx_relinked_generates_an_equivalent_matcher(tmp_path):
    """rx(1) relinked, unchanged (src/rx/main.c: per-pattern determinise + minimise + end id,
    fsm_union_array, determinise of the union -- BASELINE config 3's construction): the C matcher
    it generates must return the same (match, pattern id) for every test string as the one the
    reference's rx generates."""
    from rxcheck import STRINGS, verdicts
    got, want = verdicts(RX_CPU, tmp_path / "cpu"), verdicts(RX_REF, tmp_path / "ref")
    assert got == want and want.count(b"\n") == len(STRINGS) and b"1 5" in want


def test_eager_selftest_program():
    """libfsm_b200/shim/shim_eager_selftest.c: fsm_union_repeated_pattern_group -> determinise ->
    minimise through the shim, then fsm_exec + fsm_eager_output_cb and the additive
    fsm_exec_batch_eager must report the same, expected, id sets."""
    p = subprocess.run([os.path.join(CPU_DIR, "shim_eager_selftest")], capture_output=True, timeout=120)
    assert p.returncode == 0, (p.stdout.decode(), p.stderr.decode())
    assert b"shim eager selftest ok" in p.stdout


def test_patched_retest_impl_gpu_over_the_stub_engine():
    """integration/retest_impl_gpu.patch applied to a copy of the reference's src/retest (build time):
    `retest -l gpu` matches every vector of a regexp block with one fsm_exec_batch call and must agree,
    line for line, with the DFAVM interpreter on the reference's tests/retest/*.tst (115 vectors)."""
    import retestcheck
    if not os.path.exists(os.path.join(ROOT, "build", "shim_cpu", "retest_b200")):
        pytest.skip("patched retest not built (needs the reference tree at build time)")
    assert retestcheck.check_all(os.path.join(ROOT, "build", "shim_cpu")) >= 100
