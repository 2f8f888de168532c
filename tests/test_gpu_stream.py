"""GPU: K1b -- one long input (fsm_exec over a stream) and its shard-map form."""
import copy
import os

import numpy as np
import pytest

import goldenio
import libfsm_b200 as L
from libfsm_b200 import sharding, workloads

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


@pytest.fixture(params=["default", "256", "4096"])
def chunk_env(request):
    if request.param == "default":
        os.environ.pop("FSM_B200_STREAM_CHUNK", None)
    else:
        os.environ["FSM_B200_STREAM_CHUNK"] = request.param
    yield request.param
    os.environ.pop("FSM_B200_STREAM_CHUNK", None)


def with_start(f, s):
    g = copy.copy(f)
    g.start = s
    g._keep = []
    return g


def test_cfg2_stream_matches_oracle(oracle, torch_cuda, chunk_env):
    fsm = case("cfg2:uniform")["fsm"]
    data = workloads.cfg2_host(1, (1 << 20) + 77, adversarial=True, seed=8).reshape(-1)
    with L.Dfa(fsm) as dfa:
        for cut in (data.size, data.size - 1, 70000, 4096, 4095, 129, 5, 0):
            want = oracle.exec(fsm, data[:cut].tobytes())
            assert dfa.exec_stream(data[:cut]) == want, cut
        dev = torch_cuda.from_numpy(data).cuda()
        assert dfa.exec_stream(dev) == oracle.exec(fsm, data.tobytes())


def test_config1_digits_stream(oracle, chunk_env):
    """BASELINE config 1 shape: /[0-9]+\\.[0-9]+/ over 1 MiB of seeded ASCII (match and no-match)."""
    fsm = case("cfg1:digits")["fsm"]
    rng = np.random.default_rng(1)
    digits = rng.integers(ord("0"), ord("9") + 1, size=1 << 20, dtype=np.uint8)
    nomatch = digits.copy()
    digits[rng.random(digits.size) < 1 / 64] = ord(".")
    letters = rng.integers(ord("a"), ord("z") + 1, size=1 << 20, dtype=np.uint8)
    with L.Dfa(fsm) as dfa:
        for buf in (digits, nomatch, letters):
            want = oracle.exec(fsm, buf.tobytes())
            assert dfa.exec_stream(buf) == want
        assert oracle.exec(fsm, digits.tobytes())[0] == 1 and oracle.exec(fsm, letters.tobytes())[0] == 0


@pytest.mark.parametrize("bad_at", [None, 0, 1, 63, 64, 65, 255, 256, 300, 4095, 4096, 4097, 100000, 299999])
def test_utf8_stream_first_invalid_offset(oracle, chunk_env, bad_at):
    fsm = case("utf8:")["fsm"]
    data = workloads.utf8_host(300000, seed=5).copy()
    if bad_at is not None:
        data[bad_at] = 0xFF                                  # never valid in UTF-8
    want = oracle.exec(fsm, data.tobytes())
    with L.Dfa(fsm) as dfa:
        got = dfa.exec_stream(data)
    assert got == want
    if bad_at is not None:
        assert got[0] == 0 and got[2] <= bad_at


def test_anchored_stream_dies_in_prefix_and_body(oracle, chunk_env):
    fsm = case("anchored:^[a-f0-9]{32}$")["fsm"]
    with L.Dfa(fsm) as dfa:
        for data in (b"0123456789abcdef0123456789abcdef", b"0123456789abcdef0123456789abcdef" + b"0" * 5000,
                     b"x" + b"0" * 5000, b"0" * 31 + b"g" + b"0" * 9000):
            buf = np.frombuffer(data, dtype=np.uint8)
            assert dfa.exec_stream(buf) == oracle.exec(fsm, data)


def test_wide_table_short_input_is_exact(oracle):
    fsm = case("union6:")["fsm"]                              # > 256 states: 16-bit entries
    rng = np.random.default_rng(3)
    al = np.frombuffer(b"erowaningfld id=0123456789ABC-xy", dtype=np.uint8)
    data = al[rng.integers(0, al.size, 20000)]
    with L.Dfa(fsm) as dfa:
        assert not dfa.info["smem_resident"] or dfa.info["entry_bytes"] == 2
        assert dfa.exec_stream(data) == oracle.exec(fsm, data.tobytes())


def test_wide_class_indexed_table_stream_is_chunked_and_exact(oracle, torch_cuda):
    """K1b on a table the round-1 prefix kernel could not take (1474 states, class-indexed 16-bit rows):
    the config-3 union automaton over 8 MiB of concatenated log lines, cut at several lengths, against
    the oracle's serial walk -- and more than one launch (prefix + body + compose), i.e. not one lane."""
    c = goldenio.load_cfg3()["eager"]
    fsm = c["fsm"]
    _, inst = workloads.cfg3_patterns()
    base, _ = workloads.cfg3_lines_host(60000, inst, seed=21)
    data = base[:8 << 20]
    with L.Dfa(fsm) as dfa:
        assert dfa.info["entry_bytes"] == 2 and dfa.info["nstates"] > 1024
        for cut in (data.size, data.size - 3, 3 << 20, (1 << 21) + 1):
            L.launch_count(reset=True)
            got = dfa.exec_stream(data[:cut])
            assert L.launch_count() >= 4, "expected the chunked path"
            assert got == oracle.exec(fsm, data[:cut].tobytes()), cut
        assert dfa.exec_stream(torch_cuda.from_numpy(data).cuda()) == oracle.exec(fsm, data.tobytes())
    a = goldenio.load_cfg3()["anchored"]["fsm"]                  # dies at the first byte that fits no pattern
    with L.Dfa(a) as dfa:
        assert dfa.exec_stream(data[:4 << 20]) == oracle.exec(a, data[:4 << 20].tobytes())


@pytest.mark.parametrize("missing", [0.0, 0.002, 0.0001])
def test_wide_random_dfa_stream_every_image_survives(oracle, missing):
    """Worst case for the speculation, still exact: a random 300-state DFA (16-bit dense rows) whose chains
    from different entry states do not merge, so every chunk keeps many live images; with a few missing
    edges the walk dies somewhere in a body job."""
    import synth
    fsm = synth.dfa_from_classes(np.arange(256) % 7, 300, seed=5, missing=missing)
    rng = np.random.default_rng(6)
    data = rng.integers(0, 256, size=700000, dtype=np.uint8)
    with L.Dfa(fsm) as dfa:
        assert dfa.info["entry_bytes"] == 2
        for cut in (data.size, 400001):
            assert dfa.exec_stream(data[:cut]) == oracle.exec(fsm, data[:cut].tobytes()), (missing, cut)


@pytest.mark.parametrize("name", ["utf8:", "cfg2:uniform", "anchored:^abc[0-9]+x$"])
def test_shard_maps_equal_bruteforce_and_compose(oracle, torch_cuda, chunk_env, name):
    torch = torch_cuda
    fsm = case(name)["fsm"]
    if name == "utf8:":
        data = workloads.utf8_host(40000, seed=9).copy()
        data[33333] = 0xC0
    elif name.startswith("cfg2"):
        data = workloads.cfg2_host(1, 40000, True, seed=2).reshape(-1)
    else:
        data = np.frombuffer(b"abc" + b"7" * 30000 + b"x", dtype=np.uint8).copy()
    with L.Dfa(fsm) as dfa:
        dev = torch.from_numpy(data).cuda()
        ranges = sharding.byte_ranges(data.size, 3, align=16)
        ms, md, mf = [], [], []
        for lo, hi in ranges:
            a, b, c = dfa.exec_stream_map(dev[lo:hi])
            ms.append(a); md.append(b); mf.append(c)
            rec = torch.zeros((fsm.nstates, 2), dtype=torch.int64, device="cuda")      # asynchronous form: same records
            dfa.exec_stream_map_async(dev[lo:hi], rec)
            torch.cuda.synchronize()
            a2, b2, c2 = L.stream_map_arrays(rec.cpu().numpy())
            n = fsm.nstates
            assert (b2 == b[:n]).all() and (a2[b2 == NODEATH] == a[:n][b2 == NODEATH]).all() and (c2[b2 != NODEATH] == c[:n][b2 != NODEATH]).all()
            # brute force: the oracle from every entry state
            for s in range(0, fsm.nstates, max(1, fsm.nstates // 16)):
                ret, end, cons = oracle.exec(with_start(fsm, s), data[lo:hi].tobytes(), validate=False)
                if cons < hi - lo:
                    assert int(b[s]) == cons and int(c[s]) == end, (s, lo, hi)
                else:
                    assert int(b[s]) == 0xFFFFFFFFFFFFFFFF and int(a[s]) == end, (s, lo, hi)
        dead_row = None if dfa.info["complete"] else dfa.info["ntable_states"] - 1
        st, consumed, died = sharding.compose_stream_maps(fsm.start, dead_row, [h - l for l, h in ranges], ms, md, mf)
        ret, end, cons = oracle.exec(fsm, data.tobytes())
        assert (st, consumed) == (end, cons)
        assert (ret == 1) == (not died and bool(fsm.is_end[st]))


def test_full_size_stream_properties(torch_cuda, oracle):
    """2 GiB of valid UTF-8 resident in HBM (config 4's per-GPU share): accepted, fully
    consumed; one corrupted byte anywhere -> rejected with that exact offset."""
    torch = torch_cuda
    fsm = case("utf8:")["fsm"]
    block = workloads.utf8_host(1 << 24, seed=6)
    block = np.concatenate([block, np.full((-block.size) % 16, ord("a"), dtype=np.uint8)])   # whole code points, 16 | size
    assert oracle.exec(fsm, block.tobytes())[0] == 1
    reps = (1 << 31) // block.size
    dev = torch.from_numpy(block).cuda().repeat(reps)
    n = int(dev.numel())
    with L.Dfa(fsm) as dfa:
        ret, end, cons = dfa.exec_stream(dev)
        assert (ret, cons) == (1, n) and end == oracle.exec(fsm, block.tobytes())[1]
        for pos in (0, n // 3 + 5, n - 1):
            old = int(dev[pos]); dev[pos] = 0xFF
            ret, end, cons = dfa.exec_stream(dev)
            dev[pos] = old
            assert ret == 0 and cons <= pos and pos - cons < 4, (pos, cons)
