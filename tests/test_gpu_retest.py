"""GPU: the reference's retest(1) with the added IMPL_GPU implementation (SURVEY.md section 8(f)1;
integration/retest_impl_gpu.patch applied to a COPY of src/retest at build time, linked to the shim):
`retest -l gpu` replays the reference's own tests/retest/*.tst through fsm_determinise / fsm_minimise (K2 / K3)
and ONE fsm_exec_batch per regexp block (K1 lines kernel), and must report what the DFAVM interpreter reports."""
import os
import subprocess

import pytest

import retestcheck

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "build", "shim")


@pytest.mark.skipif(not os.path.exists(os.path.join(BUILD, "retest_b200")), reason="patched retest not built")
def test_retest_l_gpu_replays_the_reference_vectors():
    assert retestcheck.check_all(BUILD) >= 100


@pytest.mark.skipif(not os.path.exists(os.path.join(BUILD, "reperf_b200")), reason="patched reperf not built")
def test_reperf_l_gpu_runs(tmp_path):
    """reperf's driver format (src/retest/reperf.c:47-82): one pattern, a file that matches, 3 iterations."""
    text = tmp_path / "t.txt"
    text.write_bytes(b"x" * 70000 + b"12.5" + b"y" * 70000)
    scr = tmp_path / "t.scr"
    scr.write_text(f"- digits\nD pcre\nM [0-9]+\\.[0-9]+\nF {text}\nN 3\nR 1\nX\n")
    p = subprocess.run([os.path.join(BUILD, "reperf_b200"), "-l", "gpu", str(scr)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-500:] + p.stderr[-500:]
