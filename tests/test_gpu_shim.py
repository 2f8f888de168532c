"""GPU: the drop-in boundary.  The reference's re(1) relinked, UNCHANGED, against the shim
(reference libfsm minus src/libfsm/exec.c plus libfsm_b200/shim/fsm_b200_shim.c): same exit
status and output as the reference's own re(1) -- BASELINE config 1's plumbing."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RE_B200 = os.path.join(ROOT, "build", "shim", "re_b200")
RE_REF = os.path.join(ROOT, "oracle", "_ref", "re_ref")

needs_bins = pytest.mark.skipif(not (os.path.exists(RE_B200) and os.path.exists(RE_REF)),
                                reason="relinked CLIs not built (need the reference tree at build time)")


def run(binary, args):
    p = subprocess.run([binary] + args, capture_output=True, timeout=120)
    return p.returncode, p.stdout, p.stderr


@needs_bins
def test_config1_files(tmp_path):
    """re -r pcre -x '[0-9]+\\.[0-9]+' FILE over 1 MiB of seeded ASCII."""
    rng = np.random.default_rng(1)
    match = rng.integers(ord("0"), ord("9") + 1, size=1 << 20, dtype=np.uint8)
    match[rng.random(match.size) < 1 / 64] = ord(".")
    nomatch = rng.integers(ord("a"), ord("z") + 1, size=1 << 20, dtype=np.uint8)
    fm, fn = tmp_path / "match.txt", tmp_path / "nomatch.txt"
    fm.write_bytes(match.tobytes()); fn.write_bytes(nomatch.tobytes())
    for files in ([str(fm)], [str(fn)], [str(fm), str(fn)], [str(fn), str(fm)]):
        args = ["-r", "pcre", "-x", r"[0-9]+\.[0-9]+"] + files
        got, want = run(RE_B200, args), run(RE_REF, args)
        assert got[0] == want[0] and got[1] == want[1], (files, got, want)
    assert run(RE_B200, ["-r", "pcre", "-x", r"[0-9]+\.[0-9]+", str(fm)])[0] == 0
    assert run(RE_B200, ["-r", "pcre", "-x", r"[0-9]+\.[0-9]+", str(fn)])[0] == 1


@needs_bins
@pytest.mark.parametrize("args", [
    ["-r", "pcre", r"a[ -~]{7}\z", "xxabcdefgh"],
    ["-r", "pcre", r"a[ -~]{7}\z", "xxabcdefg"],
    ["-r", "pcre", r"^abc[0-9]+x$", "abc123x", "abc12", "zzz"],
    ["-r", "native", "ab*c", "abbbc"],
    ["-r", "glob", "*.txt", "notes.txt"],
    ["-r", "literal", "hello", "hello"],
])
def test_argv_strings_same_exit_status(args):
    got, want = run(RE_B200, args), run(RE_REF, args)
    assert got[0] == want[0] and got[1] == want[1], (args, got, want)


@needs_bins
def test_multi_pattern_prints_matching_pattern():
    """re -z: several patterns (end id = argv index), `match: /pattern/` printed from the end id."""
    args = ["-r", "pcre", "-z", "-y", "/dev/null"]
    # patterns come from -y FILE or -s; use the simplest form both binaries accept
    args = ["-r", "pcre", "-z", "abc", "def", "xyz"]
    got, want = run(RE_B200, args), run(RE_REF, args)
    assert got[0] == want[0] and got[1] == want[1]
