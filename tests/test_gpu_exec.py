"""GPU parity tests proper: the CUDA path (through the C ABI) against the golden fixtures
recorded from the reference, against the oracle on seeded inputs, and -- at BASELINE sizes --
through size-independent properties of the domain."""
import errno
import os
import zlib

import numpy as np
import pytest

import goldenio
import reflib
import libfsm_b200 as L
from libfsm_b200 import workloads

pytestmark = pytest.mark.gpu

CASES = goldenio.load_exec_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_exec.npz"))
BY_NAME = {c["name"]: c for c in CASES}
TILE_VARIANTS = ("tile64", "tile32", "tile128", "tile64x3")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(autouse=True)
def _reset_variant():
    L.set_exec_variant("auto")
    yield
    L.set_exec_variant("auto")


def assert_records_equal(got, want, what=""):
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, f"{what}: {bad.size} records differ, first {bad[0]}: got {got[bad[0]]} want {want[bad[0]]}"


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_host_path(case):
    """fsm_b200_exec_batch_host == the reference's recorded fsm_exec results."""
    f = case["fsm"]
    if not case["is_dfa"]:
        with pytest.raises(L.FsmB200Error) as ei:
            L.Dfa(f)
        assert ei.value.errno == errno.EINVAL          # exec.c:106-114
        return
    with L.Dfa(f) as dfa:
        for variant in ("auto", "lane"):
            L.set_exec_variant(variant)
            got = dfa.exec_batch(case["base"], case["offsets"])
            assert_records_equal(got, case["expect_amortised"], f"{case['name']} [{variant}]")
            m = case["expect"]["ret"] == 1
            assert (got["end"][m] == case["expect"]["end"][m]).all()
            assert (got["ret"] == case["expect"]["ret"]).all()


@pytest.mark.parametrize("case", [c for c in CASES if c["is_dfa"]], ids=lambda c: c["name"])
def test_golden_host_path_byte_class_tables(case, monkeypatch):
    """Same golden records through the byte-class-compressed table layout (rows indexed by
    class; normally used for DFAs too big for dense shared-memory rows), forced here."""
    monkeypatch.setenv("FSM_B200_FORCE_CLASSED", "1")
    with L.Dfa(case["fsm"]) as dfa:
        assert dfa.info["nclasses"] >= 1
        got = dfa.exec_batch(case["base"], case["offsets"])
        assert_records_equal(got, case["expect_amortised"], case["name"])
        import reflib
        assert (dfa.table() == reflib.Oracle().flatten(case["fsm"])).all()


@pytest.mark.parametrize("case", [c for c in CASES if c["is_dfa"]][::3], ids=lambda c: c["name"])
def test_device_table_matches_oracle_flatten(oracle, case):
    with L.Dfa(case["fsm"]) as dfa:
        assert (dfa.table() == oracle.flatten(case["fsm"])).all()
        assert dfa.info["nstates"] == case["fsm"].nstates


@pytest.mark.parametrize("adversarial", [False, True])
@pytest.mark.parametrize("variant", ("lane", "kstride") + TILE_VARIANTS)
def test_cfg2_fixed_stride_all_variants(torch_cuda, oracle, variant, adversarial):
    torch = torch_cuda
    fsm = BY_NAME["cfg2:uniform"]["fsm"]
    n, length = 20000 + 7, 1024                       # n not a multiple of 32: ragged last tile
    host = workloads.cfg2_host(n, length, adversarial, seed=5)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(length)
    want = oracle.exec_batch(fsm, host.reshape(-1), offsets, nthreads=8)
    with L.Dfa(fsm) as dfa:
        L.set_exec_variant(variant)
        dev = torch.from_numpy(host).cuda()
        out = dfa.exec_batch(dev, stride=length, length=length, n=n)
        torch.cuda.synchronize()
        assert_records_equal(L.results_from_torch(out), want, variant)


@pytest.mark.parametrize("variant", ("lane", "kstride") + TILE_VARIANTS)
@pytest.mark.parametrize("length,stride", [(1, 16), (15, 16), (16, 16), (17, 32), (63, 64), (64, 64), (65, 80),
                                           (127, 128), (200, 208), (1000, 1008), (4096 + 48, 4096 + 48)])
def test_odd_lengths_and_strides(torch_cuda, oracle, variant, length, stride):
    """Stage/sector remainders: lengths around every tile and vector boundary; stride > length
    leaves gap bytes that must be ignored."""
    torch = torch_cuda
    fsm = BY_NAME["cfg2:uniform"]["fsm"]
    n = 1000
    rng = np.random.default_rng(length * 131 + stride)
    host = rng.integers(0x20, 0x7F, size=(n, stride), dtype=np.uint8)
    host[rng.random((n, stride)) < 0.3] = ord("a")
    strs = [host[i, :length].tobytes() for i in range(n)]
    base, off = reflib.offsets_for(strs)
    want = oracle.exec_batch(fsm, base, off, nthreads=4)
    with L.Dfa(fsm) as dfa:
        L.set_exec_variant(variant)
        dev = torch.from_numpy(host).cuda()
        out = dfa.exec_batch(dev, stride=stride, length=length, n=n)
        torch.cuda.synchronize()
        assert_records_equal(L.results_from_torch(out), want, f"{variant} len={length} stride={stride}")


@pytest.mark.parametrize("name", ["anchored:^abc[0-9]+x$", "anchored:^[a-f0-9]{32}$", "utf8:", "union6:", "cfg1:digits"])
@pytest.mark.parametrize("variant", ("lane", "tile64", "tile32", "kstride", "auto"))
def test_dead_states_and_wide_tables_fixed_stride(torch_cuda, oracle, name, variant):
    """Incomplete DFAs (inputs die mid-way: consumed offset + stop state), the 16-bit-entry
    table (>256 states: L2-resident, lane variant only) and the UTF-8 validator."""
    torch = torch_cuda
    case = next(c for c in CASES if c["name"].startswith(name))
    fsm = case["fsm"]
    n, length = 3001, 96
    rng = np.random.default_rng(17)
    # rows: prefixes of golden inputs padded with seeded bytes from the same alphabet
    pool = case["base"] if case["base"].size else np.frombuffer(b"abc", dtype=np.uint8)
    host = pool[rng.integers(0, pool.size, size=(n, length))].astype(np.uint8)
    for i in range(0, n, 3):                         # splice in real golden inputs at row starts
        k = int(rng.integers(0, len(case["offsets"]) - 1))
        s = case["base"][int(case["offsets"][k]):int(case["offsets"][k + 1])][:length]
        host[i, :s.size] = s
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(length)
    want = oracle.exec_batch(fsm, host.reshape(-1), offsets, nthreads=4)
    with L.Dfa(fsm) as dfa:
        L.set_exec_variant(variant)
        dev = torch.from_numpy(host).cuda()
        unsupported = (variant.startswith("tile") and (not dfa.info["smem_resident"] or dfa.info["nclasses"])) or \
            (variant == "kstride" and dfa.info["kstride"] == 0)
        if unsupported:
            with pytest.raises(L.FsmB200Error) as ei:
                dfa.exec_batch(dev, stride=length, length=length, n=n)
            assert ei.value.errno == errno.ENOTSUP
            return
        out = dfa.exec_batch(dev, stride=length, length=length, n=n)
        torch.cuda.synchronize()
        assert_records_equal(L.results_from_torch(out), want, f"{name} {variant}")
        assert (want["consumed"] < length).any() or fsm.nstates > 256 or name in ("utf8:", "cfg1:digits")


def test_kstride_tables_exist_where_expected():
    # K = 4 for <= 4 byte classes; K = 2 for <= 16 classes on tables with more than 32 rows
    # (small tables stay on the one-byte kernel); none for > 256 rows
    expect = {"cfg2:uniform": 4, "cfg1:digits": 4, "anchored:^[a-f0-9]{32}$": 4, "anchored:^abc[0-9]+x$": 0,
              "endids:union6x5": 2, "union6:": 0}
    for name, k in expect.items():
        case = next(c for c in CASES if c["name"].startswith(name))
        with L.Dfa(case["fsm"]) as dfa:
            assert dfa.info["kstride"] == k, (name, dfa.info)


def test_ragged_offsets_device_path(torch_cuda, oracle):
    torch = torch_cuda
    fsm = BY_NAME["cfg2:uniform"]["fsm"]
    base, offsets = workloads.ragged_lines_host(50000, 0, 300, seed=3, alphabet=b"a" * 30 + bytes(range(0x20, 0x7F)))
    want = oracle.exec_batch(fsm, base, offsets, nthreads=8)
    with L.Dfa(fsm) as dfa:
        dbase = torch.from_numpy(base).cuda()
        doff = torch.from_numpy(offsets.astype(np.int64)).cuda()
        out = dfa.exec_batch(dbase, doff)
        torch.cuda.synchronize()
        assert_records_equal(L.results_from_torch(out), want, "ragged device")
        # unaligned base pointer: every input shifted by one byte
        shifted = torch.empty(base.size + 1, dtype=torch.uint8, device="cuda")
        shifted[1:] = dbase
        out = dfa.exec_batch(shifted[1:], doff)
        torch.cuda.synchronize()
        assert_records_equal(L.results_from_torch(out), want, "ragged device, unaligned base")
        # the host entry point chunks + pipelines: force tiny chunks
        os.environ["FSM_B200_HOST_CHUNK_MB"] = "1"
        try:
            got = dfa.exec_batch(base, offsets)
        finally:
            del os.environ["FSM_B200_HOST_CHUNK_MB"]
        assert_records_equal(got, want, "ragged host chunked")


def test_unaligned_fixed_stride_falls_back_or_refuses(torch_cuda, oracle):
    torch = torch_cuda
    fsm = BY_NAME["cfg2:uniform"]["fsm"]
    n, length = 5000, 100                               # stride 100: not a multiple of 16
    host = workloads.cfg2_host(n, length, True, seed=9)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(length)
    want = oracle.exec_batch(fsm, host.reshape(-1), offsets, nthreads=4)
    with L.Dfa(fsm) as dfa:
        dev = torch.from_numpy(host).cuda()
        out = dfa.exec_batch(dev, stride=length, length=length, n=n)      # auto -> lane
        torch.cuda.synchronize()
        assert_records_equal(L.results_from_torch(out), want, "auto on unaligned stride")
        L.set_exec_variant("tile64")
        with pytest.raises(L.FsmB200Error) as ei:
            dfa.exec_batch(dev, stride=length, length=length, n=n)
        assert ei.value.errno == errno.ENOTSUP


def test_empty_batch_and_empty_inputs(torch_cuda, oracle):
    fsm = BY_NAME["cfg2:uniform"]["fsm"]
    with L.Dfa(fsm) as dfa:
        got = dfa.exec_batch(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
        assert got.shape == (0,)
        offsets = np.zeros(1001, dtype=np.uint64)          # 1000 empty inputs
        got = dfa.exec_batch(np.zeros(0, np.uint8), offsets)
        want = oracle.exec_batch(fsm, np.zeros(0, np.uint8), offsets)
        assert_records_equal(got, want, "empty inputs")
        assert (got["consumed"] == 0).all() and (got["end"] == fsm.start).all()


def test_one_long_input_in_a_batch(torch_cuda, oracle):
    fsm = BY_NAME["cfg2:uniform"]["fsm"]
    rng = np.random.default_rng(1)
    big = rng.integers(0x20, 0x7F, size=(1 << 22) + 5, dtype=np.uint8)
    big[-8] = ord("a")
    strs = [b"a1234567", big.tobytes(), b""]
    base, off = reflib.offsets_for(strs)
    want = oracle.exec_batch(fsm, base, off)
    with L.Dfa(fsm) as dfa:
        assert_records_equal(dfa.exec_batch(base, off), want, "long input")
        assert want["ret"][1] == 1


@pytest.mark.parametrize("adversarial", [False, True])
def test_full_size_config2_properties(torch_cuda, oracle, adversarial):
    """BASELINE config 2 at full size (2^20 x 1 KiB resident in HBM).  Checked (a) on ALL
    inputs through domain properties computed independently with torch -- /a[ -~]{7}\\z/
    matches iff byte[-8] == 'a', every input is consumed entirely, and the end state is a
    function of the 'a'-mask of the last 8 bytes -- and (b) bit-exactly against the oracle
    on a 1/16 sample."""
    torch = torch_cuda
    fsm = BY_NAME["cfg2:uniform"]["fsm"]
    n, length = 1 << 20, 1024
    dev = workloads.cfg2_device(n, length, adversarial, seed=42)
    with L.Dfa(fsm) as dfa:
        results = {}
        for variant in ("lane", "tile64", "kstride", "auto"):
            L.set_exec_variant(variant)
            out = dfa.exec_batch(dev, stride=length, length=length, n=n)
            torch.cuda.synchronize()
            results[variant] = out
        assert torch.equal(results["lane"], results["tile64"])
        assert torch.equal(results["lane"], results["kstride"]) and torch.equal(results["lane"], results["auto"])
        rec = results["tile64"].view(torch.int32).reshape(n, 4)
        ret, end = rec[:, 0], rec[:, 1]
        consumed = rec[:, 2].to(torch.int64) | (rec[:, 3].to(torch.int64) << 32)
        assert bool((consumed == length).all())
        is_a = dev[:, -8:] == ord("a")
        assert torch.equal(ret == 1, is_a[:, 0])
        weights = (1 << torch.arange(7, -1, -1, device="cuda")).to(torch.int32)
        mask = (is_a.to(torch.int32) * weights).sum(dim=1)             # 0..255
        pairs = torch.unique(torch.stack([mask, end], dim=1), dim=0)
        assert pairs.shape[0] == torch.unique(mask).numel()            # end = f(mask)
        assert torch.unique(pairs[:, 1]).numel() == pairs.shape[0]     # and f is injective
        # bit-exact on a sample
        idx = torch.arange(0, n, 16, device="cuda")
        sample = dev[idx].cpu().numpy()
        offsets = np.arange(sample.shape[0] + 1, dtype=np.uint64) * np.uint64(length)
        want = oracle.exec_batch(fsm, sample.reshape(-1), offsets, nthreads=16)
        got = L.results_from_torch(results["tile64"][idx])
        assert_records_equal(got, want, "full-size sample")


RANGE_SPECS = {
    "two-low": ((0x30, 0x39), (0x2E, 0x2E)),
    "nested-low": ((0x20, 0x7E), (0x61, 0x61)),
    "low+high": ((0x30, 0x39), (0xC2, 0xDF)),
    "two-high": ((0x80, 0xBF), (0xE0, 0xEF)),
    "single": ((0x41, 0x5A), None),
    "straddle": ((0x70, 0x8F), None),
    "edges": ((0x00, 0x00), (0xFF, 0xFF)),
    "top-of-halves": ((0x7F, 0x7F), (0x80, 0x80)),
}


@pytest.mark.parametrize("missing", [0.0, 0.15])
@pytest.mark.parametrize("spec", sorted(RANGE_SPECS))
def test_alu_classified_kstride_matches_oracle(torch_cuda, oracle, monkeypatch, spec, missing):
    """k1_kstride_kernel with RNG != 0 (byte classes derived in registers from two byte ranges,
    dfa_compile.cu: find_cell_ranges) against the oracle on inputs that use ALL 256 byte values, with
    bytes concentrated around the range boundaries; the same inputs through the class-LUT form of the
    kernel (FSM_B200_KSTRIDE_LUT) and the one-byte LANE kernel must agree."""
    import synth
    torch = torch_cuda
    r0, r1 = RANGE_SPECS[spec]
    fsm = synth.dfa_from_classes(synth.classes_from_ranges(r0, r1), 61, seed=zlib.crc32(spec.encode()) & 0xFFFF, missing=missing)
    n, length = 4099, 256
    rng = np.random.default_rng(7)
    host = rng.integers(0, 256, size=(n, length), dtype=np.uint8)
    edge_bytes = []
    for r in (r0, r1):
        if r is not None:
            edge_bytes += [max(r[0] - 1, 0), r[0], r[1], min(r[1] + 1, 255), r[0] ^ 0x80, r[1] ^ 0x80]
    pick = rng.random((n, length)) < 0.6
    host[pick] = np.array(edge_bytes, dtype=np.uint8)[rng.integers(0, len(edge_bytes), size=int(pick.sum()))]
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(length)
    want = oracle.exec_batch(fsm, host.reshape(-1), offsets, nthreads=8)
    with L.Dfa(fsm) as dfa:
        assert dfa.info["kstride"] == 4 and dfa.info["krange"] in (1, 2), dfa.info
        dev = torch.from_numpy(host).cuda()
        L.set_exec_variant("kstride")
        out = dfa.exec_batch(dev, stride=length, length=length, n=n)
        torch.cuda.synchronize()
        assert_records_equal(L.results_from_torch(out), want, f"{spec} ALU-classified")
        monkeypatch.setenv("FSM_B200_KSTRIDE_LUT", "1")
        out = dfa.exec_batch(dev, stride=length, length=length, n=n)
        torch.cuda.synchronize()
        assert_records_equal(L.results_from_torch(out), want, f"{spec} class LUTs")
        monkeypatch.delenv("FSM_B200_KSTRIDE_LUT")
        L.set_exec_variant("lane")
        out = dfa.exec_batch(dev, stride=length, length=length, n=n)
        torch.cuda.synchronize()
        assert_records_equal(L.results_from_torch(out), want, f"{spec} lane")
        if missing:
            assert (want["consumed"] < length).any()


@pytest.mark.parametrize("ntiles_extra", [500, 1776 + 778])
def test_krange_tile_tail_schedule_matches_oracle(torch_cuda, oracle, monkeypatch, ntiles_extra):
    """k1_krange_tile_kernel's tail schedule (sleeping extra warps, the last two rounds dealt to all
    warps): batches of just over one / two full rounds of 148 x 12 tiles, where the host picks it, against
    the oracle and against the same kernel with the schedule switched off."""
    torch = torch_cuda
    cases = {c["name"]: c for c in goldenio.load_exec_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_exec.npz"))}
    fsm = cases["cfg2:uniform"]["fsm"]
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    n, length = (sms * 12 + ntiles_extra) * 32 - 5, 256
    host = workloads.cfg2_host(n, length, True, seed=11)
    host[::7, 100] = 0x07                                     # some inputs die mid-way
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(length)
    want = oracle.exec_batch(fsm, host.reshape(-1), offsets, nthreads=8)
    with L.Dfa(fsm) as dfa:
        dev = torch.from_numpy(host).cuda()
        L.set_exec_variant("auto")
        for tail in ("1", "0"):
            monkeypatch.setenv("FSM_B200_KRTILE_TAIL", tail)
            out = dfa.exec_batch(dev, stride=length, length=length, n=n)
            torch.cuda.synchronize()
            assert_records_equal(L.results_from_torch(out), want, f"tail schedule {tail}")
