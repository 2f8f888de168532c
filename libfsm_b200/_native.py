"""ctypes binding of libfsm_b200.so (the C ABI declared in include/fsm_b200.h).

There is no CPU fallback: if the shared library has not been built this module raises at
import, and every compute entry point fails with EIO when no sm_100 device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

from .desc import CDesc, COwnedDesc, CResult

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfsm_b200.so")

#: every symbol include/fsm_b200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = (
    "fsm_b200_abi_version", "fsm_b200_device_count", "fsm_b200_last_error",
    "fsm_b200_dfa_compile", "fsm_b200_dfa_free", "fsm_b200_dfa_info", "fsm_b200_dfa_plan", "fsm_b200_dfa_table",
    "fsm_b200_exec_batch_host", "fsm_b200_exec_batch_dev", "fsm_b200_exec_batch_dev_gather",
    "fsm_b200_wait_flags_dev", "fsm_b200_dev_alloc", "fsm_b200_dev_free", "fsm_b200_dev_zero", "fsm_b200_dev_read",
    "fsm_b200_ipc_export", "fsm_b200_ipc_open", "fsm_b200_ipc_close",
    "fsm_b200_set_exec_variant", "fsm_b200_get_exec_variant",
    "fsm_b200_exec_stream_host", "fsm_b200_exec_stream_dev", "fsm_b200_exec_stream_map_dev",
    "fsm_b200_exec_stream_map_dev_async",
    "fsm_b200_determinise", "fsm_b200_determinise_ex", "fsm_b200_desc_free", "fsm_b200_determinise_stats",
    "fsm_b200_owned_desc_eager", "fsm_b200_dfa_eager_info",
    "fsm_b200_exec_batch_eager_host", "fsm_b200_exec_batch_eager_dev",
    "fsm_b200_minimise", "fsm_b200_minimise_stats", "fsm_b200_dfavm_load",
    "fsm_b200_launch_count",
)


class CDfaInfo(C.Structure):
    _fields_ = [("nstates", C.c_uint32), ("ntable_states", C.c_uint32), ("start", C.c_uint32),
                ("entry_bytes", C.c_uint32), ("row_pitch_bytes", C.c_uint32), ("complete", C.c_uint32),
                ("smem_resident", C.c_uint32), ("device", C.c_uint32), ("table_bytes", C.c_uint64),
                ("nclasses", C.c_uint32), ("kstride", C.c_uint32), ("kclasses", C.c_uint32), ("krange", C.c_uint32),
                ("krange_lo", C.c_uint8 * 2), ("krange_hi", C.c_uint8 * 2),
                ("lines_smem", C.c_uint32), ("lines_blob_bytes", C.c_uint32), ("lines_cols", C.c_uint32),
                ("eager_ids", C.c_uint32)]


class CDetStats(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_closure", C.c_double), ("ms_expand", C.c_double),
                ("ms_intern", C.c_double), ("ms_emit", C.c_double),
                ("dfa_states", C.c_uint64), ("dfa_groups", C.c_uint64), ("rounds", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("ms_numbering", C.c_double)]


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"libfsm_b200: native library not built ({LIB_PATH} missing). "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C libfsm_b200/csrc`. "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH, use_errno=True)
    vp, u64, sz = C.c_void_p, C.c_uint64, C.c_size_t
    P = C.POINTER
    lib.fsm_b200_abi_version.restype = C.c_int
    lib.fsm_b200_device_count.restype = C.c_int
    lib.fsm_b200_last_error.restype = C.c_char_p
    lib.fsm_b200_dfa_compile.argtypes = [P(CDesc), C.c_int, P(vp)]
    lib.fsm_b200_dfa_free.argtypes = [vp]
    lib.fsm_b200_dfa_free.restype = None
    lib.fsm_b200_dfa_info.argtypes = [vp, P(CDfaInfo)]
    lib.fsm_b200_dfa_plan.argtypes = [P(CDesc), P(CDfaInfo)]
    lib.fsm_b200_dfa_table.argtypes = [vp, vp]
    lib.fsm_b200_exec_batch_host.argtypes = [vp, vp, vp, sz, vp]
    lib.fsm_b200_exec_batch_dev.argtypes = [vp, vp, vp, u64, u64, sz, vp, vp]
    lib.fsm_b200_exec_batch_dev_gather.argtypes = [vp, vp, vp, u64, u64, sz, vp, P(vp), C.c_int, C.c_int, vp, P(vp), C.c_uint32, vp]
    lib.fsm_b200_wait_flags_dev.argtypes = [C.c_int, vp, C.c_uint32, C.c_uint32, vp, vp]
    lib.fsm_b200_dev_alloc.argtypes = [C.c_int, sz, P(vp)]
    lib.fsm_b200_dev_free.argtypes = [C.c_int, vp]
    lib.fsm_b200_dev_zero.argtypes = [C.c_int, vp, sz]
    lib.fsm_b200_dev_read.argtypes = [C.c_int, vp, vp, sz]
    lib.fsm_b200_ipc_export.argtypes = [vp, vp]
    lib.fsm_b200_ipc_open.argtypes = [C.c_int, vp, P(vp)]
    lib.fsm_b200_ipc_close.argtypes = [C.c_int, vp]
    lib.fsm_b200_set_exec_variant.argtypes = [C.c_int]
    lib.fsm_b200_get_exec_variant.restype = C.c_int
    lib.fsm_b200_exec_stream_host.argtypes = [vp, vp, u64, P(CResult)]
    lib.fsm_b200_exec_stream_dev.argtypes = [vp, vp, u64, P(CResult), vp]
    lib.fsm_b200_exec_stream_map_dev.argtypes = [vp, vp, u64, vp, vp, vp, vp]
    lib.fsm_b200_exec_stream_map_dev_async.argtypes = [vp, vp, u64, vp, vp]
    lib.fsm_b200_exec_stream_map_dev_async.restype = C.c_int
    lib.fsm_b200_determinise.argtypes = [P(CDesc), C.c_int, sz, P(COwnedDesc)]
    lib.fsm_b200_determinise_ex.argtypes = [P(CDesc), C.c_int, sz, C.c_uint, P(COwnedDesc)]
    lib.fsm_b200_owned_desc_eager.argtypes = [P(COwnedDesc), P(C.c_void_p), P(C.c_void_p)]
    lib.fsm_b200_dfa_eager_info.argtypes = [C.c_void_p, P(C.c_uint32), P(C.c_void_p)]
    lib.fsm_b200_exec_batch_eager_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.fsm_b200_exec_batch_eager_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_size_t,
                                                  C.c_void_p, C.c_void_p, C.c_void_p]
    lib.fsm_b200_desc_free.argtypes = [P(COwnedDesc)]
    lib.fsm_b200_desc_free.restype = None
    lib.fsm_b200_determinise_stats.argtypes = [P(CDetStats)]
    lib.fsm_b200_minimise.argtypes = [P(CDesc), C.c_int, P(COwnedDesc)]
    lib.fsm_b200_minimise_stats.argtypes = [P(CDetStats)]
    lib.fsm_b200_dfavm_load.argtypes = [vp, sz, P(COwnedDesc)]
    lib.fsm_b200_launch_count.argtypes = [C.c_int]
    lib.fsm_b200_launch_count.restype = C.c_uint64
    return lib


lib = _load()


class FsmB200Error(OSError):
    """A C-ABI call returned -1; carries errno and the library's error text."""


def check(rc: int, what: str) -> None:
    if rc < 0:
        err = C.get_errno()
        msg = lib.fsm_b200_last_error().decode("utf-8", "replace")
        raise FsmB200Error(err, f"{what}: {msg or os.strerror(err)}")
