"""GPU: K1b's fused form for small automata (libfsm_b200/csrc/k1b_rep.cuh: per-lane replicated table,
prefix + body + warp fold in one kernel) against the oracle's serial walk (src/libfsm/exec.c:132-151),
from EVERY entry state, at every alignment and around every length the kernel treats specially."""
import copy
import os

import numpy as np
import pytest

import goldenio
import libfsm_b200 as L
import synth
from libfsm_b200 import workloads
from libfsm_b200.desc import FlatFsm

pytestmark = pytest.mark.gpu

CASES = goldenio.load_exec_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_exec.npz"))
NODEATH = 0xFFFFFFFFFFFFFFFF


def case(prefix):
    return next(c for c in CASES if c["name"].startswith(prefix))


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def with_start(f, s):
    g = copy.copy(f)
    g.start = s
    g._keep = []
    return g


def rotation_dfa(n, missing_from=None):
    """byte b moves state s to (s + b % n) % n: chains from different entry states NEVER merge, so every
    chunk keeps n live images (n body walks per lane).  missing_from: that state has no edge on 0xFE."""
    edges = []
    for s in range(n):
        by_dst = {}
        for b in range(256):
            if missing_from == s and b == 0xFE:
                continue
            by_dst.setdefault((s + b % n) % n, []).append(b)
        edges += [(s, syms, d) for d, syms in by_dst.items()]
    return FlatFsm.from_edges(n, 0, [n - 1], edges)


def check_all_entries(oracle, dfa, fsm, dev, data):
    """exec_stream from the start state and the shard map from every entry state == the oracle's walk."""
    assert dfa.exec_stream(dev) == oracle.exec(fsm, data.tobytes())
    if data.size == 0:
        return
    ms, md, mf = dfa.exec_stream_map(dev)
    # the asynchronous form leaves the same records on the device
    import torch
    rec = torch.zeros((fsm.nstates, 2), dtype=torch.int64, device=dev.device)
    dfa.exec_stream_map_async(dev, rec)
    torch.cuda.synchronize()
    a_s, a_d, a_f = L.stream_map_arrays(rec.cpu().numpy())
    n = fsm.nstates
    assert (a_d == md[:n]).all() and (a_s[a_d == NODEATH] == ms[:n][a_d == NODEATH]).all() and \
        (a_f[a_d != NODEATH] == mf[:n][a_d != NODEATH]).all()
    for s in range(fsm.nstates):
        ret, end, cons = oracle.exec(with_start(fsm, s), data.tobytes(), validate=False)
        if cons < data.size:
            assert (int(md[s]), int(mf[s])) == (cons, end), (s, data.size)
        else:
            assert int(md[s]) == NODEATH and int(ms[s]) == end, (s, data.size)


LENGTHS = [0, 1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 129, 255, 256, 257, 575, 576, 577, 608, 1000,
           4096, 20001, 300007]


def test_small_automaton_takes_the_fused_form(oracle, torch_cuda):
    fsm = case("utf8:")["fsm"]
    data = workloads.utf8_host(1 << 20, seed=3)
    with L.Dfa(fsm) as dfa:
        assert dfa.info["ntable_states"] <= 12
        dev = torch_cuda.from_numpy(data).cuda()
        L.launch_count(reset=True)
        got = dfa.exec_stream(dev)
        assert L.launch_count() == 2, "walk + final fold"
        assert got == oracle.exec(fsm, data.tobytes())
        os.environ["FSM_B200_STREAM_REP"] = "0"
        try:
            L.launch_count(reset=True)
            assert dfa.exec_stream(dev) == got
            assert L.launch_count() >= 4, "the generic chunked path"
        finally:
            os.environ.pop("FSM_B200_STREAM_REP", None)


@pytest.mark.parametrize("off", [0, 1, 7, 16, 31])
def test_utf8_every_length_and_alignment(oracle, torch_cuda, off):
    fsm = case("utf8:")["fsm"]
    text = workloads.utf8_host(400000, seed=11)
    big = torch_cuda.from_numpy(np.concatenate([np.zeros(off, np.uint8), text])).cuda()
    with L.Dfa(fsm) as dfa:
        for n in LENGTHS:
            check_all_entries(oracle, dfa, fsm, big[off:off + n], text[:n])


@pytest.mark.parametrize("bad_at", [0, 1, 30, 31, 32, 63, 64, 65, 100, 575, 576, 577, 639, 640, 641, 4000, 123457, 299999])
@pytest.mark.parametrize("off", [0, 5])
def test_utf8_death_offsets(oracle, torch_cuda, bad_at, off):
    """0xFF never appears in UTF-8: the walk from the start state dies exactly there -- in a prefix window,
    in the head / a sector / the tail of a body walk, in the first and in the last chunk."""
    fsm = case("utf8:")["fsm"]
    text = workloads.utf8_host(300000, seed=5).copy()
    text[bad_at] = 0xFF
    big = torch_cuda.from_numpy(np.concatenate([np.zeros(off, np.uint8), text])).cuda()
    with L.Dfa(fsm) as dfa:
        for n in (text.size, bad_at + 1, bad_at + 40):
            n = min(n, text.size)
            check_all_entries(oracle, dfa, fsm, big[off:off + n], text[:n])


@pytest.mark.parametrize("nstates,missing", [(1, 0.0), (2, 0.0), (3, 0.02), (5, 0.0), (7, 0.001), (8, 0.0), (11, 0.0005), (11, 0.0), (12, 0.0)])
def test_random_small_dfas(oracle, torch_cuda, nstates, missing):
    """Seeded random automata of every size the form takes (12 rows: 11 states + dead row, or 12 complete)."""
    fsm = synth.dfa_from_classes(np.arange(256) % 5, nstates, seed=40 + nstates, missing=missing)
    rng = np.random.default_rng(nstates)
    data = rng.integers(0, 256, size=200000, dtype=np.uint8)
    dev = torch_cuda.from_numpy(data).cuda()
    with L.Dfa(fsm) as dfa:
        assert dfa.info["ntable_states"] <= 12
        for lo, n in ((0, data.size), (3, 70001), (64, 1234), (96, 577), (1, 63)):
            check_all_entries(oracle, dfa, fsm, dev[lo:lo + n], data[lo:lo + n])


@pytest.mark.parametrize("n,missing_from", [(3, None), (7, None), (11, None), (7, 4), (11, 0)])
def test_chains_that_never_merge(oracle, torch_cuda, n, missing_from):
    """Rotation automata: n distinct live images per chunk, n body walks per lane -- slower, still exact;
    with one missing edge the walks die in different chunks for different entry states."""
    fsm = rotation_dfa(n, missing_from)
    rng = np.random.default_rng(n)
    data = rng.integers(0, 253, size=150000, dtype=np.uint8)
    if missing_from is not None:
        data[[40, 700, 9000, 90001, 149999]] = 0xFE
    dev = torch_cuda.from_numpy(data).cuda()
    with L.Dfa(fsm) as dfa:
        for lo, m in ((0, data.size), (2, 100000), (0, 8000), (5, 600)):
            check_all_entries(oracle, dfa, fsm, dev[lo:lo + m], data[lo:lo + m])


def test_absorbing_states_need_no_walk(oracle, torch_cuda):
    """/x/ unanchored as a 2-state DFA: the accept state loops on every byte (absorbing image: no body walk)."""
    edges = [(0, [b for b in range(256) if b != ord("x")], 0), (0, [ord("x")], 1), (1, list(range(256)), 1)]
    fsm = FlatFsm.from_edges(2, 0, [1], edges)
    data = np.full(500000, ord("a"), dtype=np.uint8)
    with L.Dfa(fsm) as dfa:
        dev = torch_cuda.from_numpy(data).cuda()
        check_all_entries(oracle, dfa, fsm, dev, data)
        data[333333] = ord("x")
        dev = torch_cuda.from_numpy(data).cuda()
        check_all_entries(oracle, dfa, fsm, dev, data)


def test_both_forms_agree_on_a_large_input(torch_cuda):
    """256 MiB of UTF-8 with one bad byte: the fused form and the generic chunked path give the same record."""
    fsm = case("utf8:")["fsm"]
    block = workloads.utf8_host(1 << 24, seed=6)
    dev = torch_cuda.from_numpy(block).cuda().repeat(16)
    n = int(dev.numel())
    with L.Dfa(fsm) as dfa:
        for pos in (None, 17, n // 2 + 3, n - 2):
            old = None
            if pos is not None:
                old = int(dev[pos]); dev[pos] = 0xFF
            got = dfa.exec_stream(dev)
            os.environ["FSM_B200_STREAM_REP"] = "0"
            try:
                want = dfa.exec_stream(dev)
            finally:
                os.environ.pop("FSM_B200_STREAM_REP", None)
            if pos is not None:
                dev[pos] = old
                assert got[0] == 0 and got[2] <= pos and pos - got[2] < 4
            else:
                assert got[0] == 1 and got[2] == n
            assert got == want, pos


def test_tma_tile_form_is_exact(oracle, torch_cuda):
    """The opt-in TMA-tile form of the same kernel (k1b_rep_tma.cuh: compact table, input by 2-D TMA tiles;
    measured slower, kept selectable): same records from every entry state, full and partial last warps, deaths."""
    fsm = case("utf8:")["fsm"]
    text = workloads.utf8_host(3000000, seed=13).copy()
    os.environ["FSM_B200_REP_TMA"] = "1"
    try:
        with L.Dfa(fsm) as dfa:
            dev = torch_cuda.from_numpy(text).cuda()
            for n in (text.size, 2000003, 576 * 40, 576 * 32 + 5):
                check_all_entries(oracle, dfa, fsm, dev[:n], text[:n])
            for bad in (100, 70000, 2999990):
                t2 = text.copy(); t2[bad] = 0xFF
                check_all_entries(oracle, dfa, fsm, torch_cuda.from_numpy(t2).cuda(), t2)
            for stages in ("2", "3"):
                os.environ["FSM_B200_REP_TMA_STAGES"] = stages
                assert dfa.exec_stream(dev) == oracle.exec(fsm, text.tobytes())
    finally:
        os.environ.pop("FSM_B200_REP_TMA", None)
        os.environ.pop("FSM_B200_REP_TMA_STAGES", None)
