"""Shared by the CPU (stub engine) and GPU tests of the patched retest(1): `retest -l gpu` (IMPL_GPU,
integration/retest_impl_gpu.patch: libfsm's own fsm_exec_batch, one call per regexp block) must report,
line for line, what the reference's default implementation (`-l vm`, the DFAVM interpreter built by the
same binary) reports on the reference's own tests/retest/*.tst vectors."""
import glob
import os
import subprocess


def run_retest(binary: str, tst: str, impl: str | None):
    cmd = [binary] + (["-l", impl] if impl else []) + [tst]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    lines = [l for l in p.stdout.splitlines() if l.startswith("[") and not l.startswith("[TIME")]
    return p.returncode, lines, p.stdout[-400:] + p.stderr[-400:]


def check_all(build_dir: str):
    binary = os.path.join(build_dir, "retest_b200")
    tsts = sorted(glob.glob(os.path.join(build_dir, "retest_tst", "*.tst")))
    assert len(tsts) >= 5
    total = 0
    for t in tsts:
        rc_g, gpu, tail = run_retest(binary, t, "gpu")
        rc_v, vm, _ = run_retest(binary, t, None)
        assert rc_g == 0 and rc_v == 0, (t, rc_g, rc_v, tail)
        assert gpu == vm, (t, [x for x in zip(gpu, vm) if x[0] != x[1]][:3])
        assert all(l.startswith("[OK    ]") for l in gpu), t
        total += len(gpu)
    assert total >= 100
    return total
