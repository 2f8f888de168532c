"""CPU: the C-ABI library loads and exports every symbol include/fsm_b200.h declares; the
host-side logic (flat description, validation error path) works without a GPU."""
import ctypes
import errno
import os
import re

import numpy as np
import pytest

import libfsm_b200 as L
from libfsm_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "fsm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fsm_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 15
    lib = ctypes.CDLL(L.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/fsm_b200.h but not exported"
    assert set(syms) == set(L.ABI_SYMBOLS)


def test_abi_version():
    assert _native.lib.fsm_b200_abi_version() == 4


def test_struct_layouts():
    from libfsm_b200.desc import CDesc, CResult
    assert ctypes.sizeof(CResult) == 16
    assert ctypes.sizeof(CDesc) == 16 + 8 * 8


def test_no_cpu_fallback_without_gpu():
    if L.device_count() > 0:
        pytest.skip("a GPU is present")
    f = L.FlatFsm.from_edges(2, 0, [1], [(0, ord("a"), 1), (1, range(256), 1)])
    with pytest.raises(L.FsmB200Error) as ei:
        L.Dfa(f)
    assert ei.value.errno == errno.EIO


def test_not_a_dfa_is_einval_before_touching_the_device():
    # validation (exec.c:106-114 semantics) happens on the host, so this works without a GPU
    nfa = L.FlatFsm.from_edges(3, 0, [2], [(0, ord("a"), 1), (0, ord("a"), 2)])
    with pytest.raises(L.FsmB200Error) as ei:
        L.Dfa(nfa)
    assert ei.value.errno == errno.EINVAL
    eps = L.FlatFsm.from_edges(2, 0, [1], [(0, ord("a"), 1)], eps=[(0, 1)])
    with pytest.raises(L.FsmB200Error) as ei:
        L.Dfa(eps)
    assert ei.value.errno == errno.EINVAL
    nostart = L.FlatFsm.from_edges(2, None, [1], [(0, ord("a"), 1)])
    with pytest.raises(L.FsmB200Error) as ei:
        L.Dfa(nostart)
    assert ei.value.errno == errno.EINVAL


def test_flatfsm_roundtrip_npz(tmp_path):
    f = L.FlatFsm.from_edges(3, 0, [2], [(0, range(97, 123), 1), (1, ord("x"), 2)], endids={2: [7, 3]})
    p = str(tmp_path / "f.npz")
    f.save(p)
    g = L.FlatFsm.load(p)
    assert (g.dense_table() == f.dense_table()).all()
    assert list(g.endids_of(2)) == [3, 7]
