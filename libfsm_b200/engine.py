"""Host-side mirror of the reference interface for the hot path.

``Dfa`` plays the role of a determinised ``struct fsm`` handed to ``fsm_exec``
(reference include/fsm/fsm.h:560-562): it is compiled once (DFA-ness validated as
src/libfsm/exec.c:106-114 does per call) and then executes batches of inputs on the GPU.
``determinise`` mirrors ``fsm_determinise_with_config`` (include/fsm/fsm.h:472-488).

numpy arrays are host buffers (end-to-end path, copies inside the call); torch CUDA
tensors are device buffers (resident path, asynchronous on the current stream).
"""
from __future__ import annotations

import ctypes as C
import errno as _errno

import numpy as np

from . import _native
from ._native import lib, check, CDfaInfo, CDetStats
from .desc import FlatFsm, COwnedDesc, CResult, RESULT_DTYPE

VARIANTS = {"auto": 0, "lane": 1, "tile64": 2, "tile32": 3, "tile128": 4, "tile64x3": 5, "kstride": 6}


def device_count() -> int:
    return int(lib.fsm_b200_device_count())


def set_exec_variant(name_or_id) -> None:
    v = VARIANTS[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
    check(lib.fsm_b200_set_exec_variant(v), "set_exec_variant")


def launch_count(reset: bool = False) -> int:
    return int(lib.fsm_b200_launch_count(1 if reset else 0))


def plan(fsm: FlatFsm) -> dict:
    """The table layout a compile would choose for `fsm`, computed on the host (no device needed):
    entry width, row pitch, shared-memory residency, byte classes, k-stride.  Raises like
    ``Dfa(...)`` (EINVAL) when `fsm` is not a DFA."""
    info = CDfaInfo()
    cdesc = fsm.as_c()
    check(lib.fsm_b200_dfa_plan(C.byref(cdesc), C.byref(info)), "dfa_plan")
    return _info_dict(info)


def _info_dict(info) -> dict:
    out = {}
    for k, _ in CDfaInfo._fields_:
        v = getattr(info, k)
        out[k] = int(v) if isinstance(v, int) else [int(x) for x in v]
    return out


def _is_torch_cuda(x) -> bool:
    return hasattr(x, "is_cuda") and bool(x.is_cuda)


def _stream_ptr() -> int:
    import torch
    return int(torch.cuda.current_stream().cuda_stream)


class Dfa:
    """A compiled, device-resident DFA (``fsm_b200_dfa``)."""

    def __init__(self, fsm: FlatFsm, device: int = 0):
        self.fsm = fsm
        self.device = device
        self._h = C.c_void_p()
        cdesc = fsm.as_c()
        check(lib.fsm_b200_dfa_compile(C.byref(cdesc), device, C.byref(self._h)), "dfa_compile")
        info = CDfaInfo()
        check(lib.fsm_b200_dfa_info(self._h, C.byref(info)), "dfa_info")
        self.info = _info_dict(info)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            lib.fsm_b200_dfa_free(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def table(self) -> np.ndarray:
        """Dense [nstates, 256] table as read back from the device (0xFFFFFFFF = no edge)."""
        out = np.empty((self.info["nstates"], 256), dtype=np.uint32)
        check(lib.fsm_b200_dfa_table(self._h, out.ctypes.data), "dfa_table")
        return out

    # ---- eager outputs ----------------------------------------------------------------
    def eager_ids(self) -> np.ndarray:
        """The distinct eager-output ids of this DFA, ascending: bit b of a mask <=> eager_ids()[b]."""
        nbits, ids = C.c_uint32(0), C.c_void_p()
        check(lib.fsm_b200_dfa_eager_info(self._h, C.byref(nbits), C.byref(ids)), "dfa_eager_info")
        if nbits.value == 0:
            return np.zeros(0, np.uint32)
        return np.ctypeslib.as_array(C.cast(ids, C.POINTER(C.c_uint32)), shape=(nbits.value,)).copy()

    def exec_batch_eager(self, base, offsets=None, *, stride=None, length=None, n=None):
        """(records, masks uint64 [n, words]): the usual records plus, per input, the bitset of
        eager-output ids fired along its walk (``fired_ids`` decodes one).  Host: numpy ``base`` /
        ``offsets``.  Device: torch.uint8 CUDA ``base`` with torch int64 CUDA ``offsets`` or a fixed
        ``stride``/``length``/``n``; returns torch tensors ([n, 16] uint8 and [n, words] int64)."""
        if _is_torch_cuda(base):
            import torch
            words = (len(self.eager_ids()) + 63) // 64
            if offsets is not None:
                n = int(offsets.numel()) - 1
                off_ptr, stride, length = offsets.data_ptr(), 0, 0
            else:
                off_ptr = None
                length = stride if length is None else length
                n = int(base.numel()) // int(stride) if n is None else n
            out = torch.empty((max(n, 0), 16), dtype=torch.uint8, device=base.device)
            masks = torch.zeros((max(n, 0), max(words, 1)), dtype=torch.int64, device=base.device)
            check(lib.fsm_b200_exec_batch_eager_dev(self._h, base.data_ptr(), off_ptr, int(stride), int(length), n,
                                                    out.data_ptr(), masks.data_ptr(), _stream_ptr()), "exec_batch_eager_dev")
            return out, masks
        base = np.ascontiguousarray(base, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.shape[0] - 1
        words = (len(self.eager_ids()) + 63) // 64
        out = np.empty(max(n, 0), dtype=RESULT_DTYPE)
        masks = np.zeros((max(n, 0), max(words, 1)), dtype=np.uint64)
        bptr = base.ctypes.data if base.size else np.zeros(16, np.uint8).ctypes.data
        check(lib.fsm_b200_exec_batch_eager_host(self._h, bptr, offsets.ctypes.data, n, out.ctypes.data, masks.ctypes.data),
              "exec_batch_eager_host")
        return out, masks[:, :words]

    def fired_ids(self, mask_row) -> list:
        ids = self.eager_ids()
        return [int(ids[b]) for b in range(len(ids)) if (int(mask_row[b >> 6]) >> (b & 63)) & 1]

    # ---- batched execution -----------------------------------------------------------
    def exec_batch(self, base, offsets=None, *, stride=None, length=None, n=None, out=None):
        """n independent fsm_exec calls.

        host:   ``base`` uint8 numpy array, ``offsets`` uint64 numpy array [n+1]
                -> numpy structured array (ret, end, consumed)
        device: ``base`` torch.uint8 CUDA tensor; either ``offsets`` (torch int64 CUDA
                tensor [n+1]) or fixed ``stride``/``length``/``n``
                -> torch.uint8 CUDA tensor [n, 16] (view with ``results_from_torch``)
        """
        if _is_torch_cuda(base):
            import torch
            if offsets is not None:
                n = int(offsets.numel()) - 1
                off_ptr, stride, length = offsets.data_ptr(), 0, 0
            else:
                off_ptr = None
                if length is None:
                    length = stride
                if n is None:
                    n = int(base.numel()) // int(stride) if stride else 0
            if out is None:
                out = torch.empty((max(n, 0), 16), dtype=torch.uint8, device=base.device)
            check(lib.fsm_b200_exec_batch_dev(self._h, base.data_ptr(), off_ptr, int(stride), int(length),
                                              n, out.data_ptr(), _stream_ptr()), "exec_batch_dev")
            return out
        base = np.ascontiguousarray(base, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.shape[0] - 1
        if out is None:
            out = np.empty(max(n, 0), dtype=RESULT_DTYPE)
        bptr = base.ctypes.data if base.size else np.zeros(16, np.uint8).ctypes.data
        check(lib.fsm_b200_exec_batch_host(self._h, bptr, offsets.ctypes.data, n, out.ctypes.data),
              "exec_batch_host")
        return out

    def exec_batch_gather(self, base, *, stride: int, length: int, n: int, out_ptr: int, peer_ptrs, npeers: int,
                          compact: bool = False, sig_counter: int | None = None, sig_flags=None, sig_value: int = 0) -> None:
        """Fixed-stride device batch whose records go to out_ptr AND to npeers peer buffers
        (fused scan + gather over NVLink peer memory; see libfsm_b200.peer.GatherRing).
        compact: peers receive 4-byte match ids ((ret == 1) << 31 | end) instead of records.
        sig_*: optional completion flags written into every peer's memory by the last CTA."""
        check(lib.fsm_b200_exec_batch_dev_gather(self._h, base.data_ptr(), None, int(stride), int(length), n,
                                                 out_ptr, peer_ptrs, npeers, 1 if compact else 0,
                                                 sig_counter, sig_flags, int(sig_value) & 0xFFFFFFFF, _stream_ptr()),
              "exec_batch_dev_gather")

    def exec_batch_hostptr(self, base_ptr: int, offsets_ptr: int, n: int, out_ptr: int) -> None:
        """Raw-pointer form of the host path (pinned buffers owned by the caller)."""
        check(lib.fsm_b200_exec_batch_host(self._h, base_ptr, offsets_ptr, n, out_ptr), "exec_batch_host")

    # ---- one long input -----------------------------------------------------------------
    def exec_stream(self, buf):
        """One fsm_exec call over a single long input -> (ret, end, consumed)."""
        r = CResult()
        if _is_torch_cuda(buf):
            check(lib.fsm_b200_exec_stream_dev(self._h, buf.data_ptr(), int(buf.numel()), C.byref(r),
                                               _stream_ptr()), "exec_stream_dev")
        else:
            buf = np.ascontiguousarray(buf, dtype=np.uint8)
            ptr = buf.ctypes.data if buf.size else np.zeros(16, np.uint8).ctypes.data
            check(lib.fsm_b200_exec_stream_host(self._h, ptr, int(buf.size), C.byref(r)), "exec_stream_host")
        return int(r.ret), int(r.end), int(r.consumed)

    def exec_stream_map(self, dbuf):
        """Shard form: per entry state the (exit state, first dead offset, dead-from state)
        of the byte range ``dbuf`` (torch CUDA uint8).  See include/fsm_b200.h."""
        nt = self.info["ntable_states"]
        ms = np.empty(nt, np.uint32); md = np.empty(nt, np.uint64); mf = np.empty(nt, np.uint32)
        check(lib.fsm_b200_exec_stream_map_dev(self._h, dbuf.data_ptr(), int(dbuf.numel()),
                                               ms.ctypes.data, md.ctypes.data, mf.ctypes.data,
                                               _stream_ptr()), "exec_stream_map_dev")
        return ms, md, mf

    def exec_stream_map_async(self, dbuf, out):
        """The shard map left on the device: ``out`` (torch CUDA int64 [nstates, 2]) receives one
        fsm_b200_stream_map_entry per entry state, written by work queued on the current stream; nothing
        waits.  Decode a (gathered) copy with stream_map_arrays()."""
        assert out.is_cuda and out.is_contiguous() and out.numel() * out.element_size() >= 16 * self.info["nstates"]
        check(lib.fsm_b200_exec_stream_map_dev_async(self._h, dbuf.data_ptr(), int(dbuf.numel()), out.data_ptr(),
                                                     _stream_ptr()), "exec_stream_map_dev_async")


def stream_map_arrays(records: np.ndarray):
    """[..., nstates, 2] int64/uint64 view of fsm_b200_stream_map_entry records -> (map_state u32, map_dead u64,
    map_dead_state u32) as Dfa.exec_stream_map returns them (minus the dead row)."""
    a = np.ascontiguousarray(records).view(np.uint64)
    state = (a[..., 0] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    died = (a[..., 0] >> np.uint64(32)) != 0
    dead = np.where(died, a[..., 1], np.uint64(0xFFFFFFFFFFFFFFFF))
    return state, dead, np.where(died, state, np.uint32(0xFFFFFFFF)).astype(np.uint32)


def results_from_torch(t) -> np.ndarray:
    """[n,16] uint8 CUDA/CPU tensor of result records -> numpy structured array."""
    return t.detach().cpu().numpy().reshape(-1).view(RESULT_DTYPE)


class StateLimitReached(Exception):
    """FSM_DETERMINISE_WITH_CONFIG_STATE_LIMIT_REACHED (reference include/fsm/fsm.h:481-488)."""


DET_REFERENCE_NUMBERING = 1


def determinise(nfa: FlatFsm, device: int = 0, state_limit: int = 0, numbering: str | None = None) -> FlatFsm:
    """GPU subset construction.  numbering=None: the library default (BFS order, i.e. the
    reference's DFA up to state renumbering, unless FSM_B200_DET_NUMBERING=reference);
    "reference": state for state the DFA fsm_determinise builds; "bfs": BFS order."""
    od = COwnedDesc()
    cdesc = nfa.as_c()
    if numbering is None:
        rc = lib.fsm_b200_determinise(C.byref(cdesc), device, state_limit, C.byref(od))
    else:
        flags = {"reference": DET_REFERENCE_NUMBERING, "bfs": 0}[numbering]
        rc = lib.fsm_b200_determinise_ex(C.byref(cdesc), device, state_limit, flags, C.byref(od))
    if rc == 1:
        raise StateLimitReached()
    check(rc, "determinise")
    try:
        return _with_eager(FlatFsm.from_c(od.desc), od)
    finally:
        lib.fsm_b200_desc_free(C.byref(od))


def minimise(dfa: FlatFsm, device: int = 0) -> FlatFsm:
    """GPU minimisation (mirrors ``fsm_minimise``, reference include/fsm/fsm.h:502-503): the
    minimal DFA, unique up to state numbering; 0 states if nothing can match."""
    od = COwnedDesc()
    cdesc = dfa.as_c()
    check(lib.fsm_b200_minimise(C.byref(cdesc), device, C.byref(od)), "minimise")
    try:
        if od.desc.nstates == 0:
            return FlatFsm(0, 0, False, np.zeros(0, np.uint8), np.zeros(1, np.uint64),
                           np.zeros((0, 4), np.uint64), np.zeros(0, np.uint32), None, None, None, None)
        return _with_eager(FlatFsm.from_c(od.desc), od)
    finally:
        lib.fsm_b200_desc_free(C.byref(od))


def load_dfavm(image: bytes) -> FlatFsm:
    """A DFA from the reference's DFAVM bytecode image ("DFAVM$", src/libfsm/vm.c:39-71; host code only):
    state i = the i-th FETCH; yes / no semantics only (no end ids in the image)."""
    od = COwnedDesc()
    buf = (C.c_uint8 * len(image)).from_buffer_copy(image)
    check(lib.fsm_b200_dfavm_load(buf, len(image), C.byref(od)), "dfavm_load")
    try:
        return FlatFsm.from_c(od.desc)
    finally:
        lib.fsm_b200_desc_free(C.byref(od))


def _with_eager(f: FlatFsm, od) -> FlatFsm:
    """Attach the eager-output sets of a library-owned result (fsm_b200_owned_desc_eager)."""
    off, ids = C.c_void_p(), C.c_void_p()
    check(lib.fsm_b200_owned_desc_eager(C.byref(od), C.byref(off), C.byref(ids)), "owned_desc_eager")
    if off and f.nstates > 0:
        eo = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_uint64)), shape=(f.nstates + 1,)).copy()
        if int(eo[-1]) > 0:
            f.eager_off = eo
            f.eager_ids = np.ctypeslib.as_array(C.cast(ids, C.POINTER(C.c_uint32)), shape=(int(eo[-1]),)).copy()
    return f


def minimise_stats() -> dict:
    st = CDetStats()
    lib.fsm_b200_minimise_stats(C.byref(st))
    return {k: getattr(st, k) for k, _ in CDetStats._fields_}


def determinise_stats() -> dict:
    st = CDetStats()
    lib.fsm_b200_determinise_stats(C.byref(st))
    return {k: getattr(st, k) for k, _ in CDetStats._fields_}
