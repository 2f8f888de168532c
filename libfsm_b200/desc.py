"""Flat automaton description shared with the C ABI (``struct fsm_b200_desc``).

Mirrors the reference's data model: per-state edge groups of (256-bit label set,
destination) exactly as ``struct edge_group`` stores them (reference
src/adt/edgeset.c:34-41), epsilon sets (src/libfsm/internal.h:52-54), end bits and
per-end-state sorted end-id sets (src/libfsm/endids.c:686-755).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np


class CDesc(C.Structure):
    """``struct fsm_b200_desc`` of include/fsm_b200.h."""
    _fields_ = [
        ("nstates", C.c_uint32), ("start", C.c_uint32), ("hasstart", C.c_uint32), ("reserved", C.c_uint32),
        ("is_end", C.c_void_p), ("group_off", C.c_void_p), ("group_symbols", C.c_void_p),
        ("group_to", C.c_void_p), ("eps_off", C.c_void_p), ("eps_to", C.c_void_p),
        ("endid_off", C.c_void_p), ("endids", C.c_void_p),
    ]


DESC_EAGER = 1


class CDescExt(C.Structure):
    """``struct fsm_b200_desc_ext``: a desc with FSM_B200_DESC_EAGER set in ``reserved`` plus the
    eager-output CSR."""
    _fields_ = [("base", CDesc), ("eager_off", C.c_void_p), ("eager_ids", C.c_void_p)]


class COwnedDesc(C.Structure):
    """``struct fsm_b200_owned_desc``."""
    _fields_ = [("desc", CDesc), ("owner", C.c_void_p)]


class CResult(C.Structure):
    """``struct fsm_b200_result`` (16 bytes)."""
    _fields_ = [("ret", C.c_int32), ("end", C.c_uint32), ("consumed", C.c_uint64)]


RESULT_DTYPE = np.dtype([("ret", "<i4"), ("end", "<u4"), ("consumed", "<u8")])
assert RESULT_DTYPE.itemsize == C.sizeof(CResult) == 16


def _arr(x, dtype, n=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=dtype))
    if n is not None and a.size == 0:
        a = np.zeros(max(n, 1), dtype=dtype)[:0]
    return a


@dataclass
class FlatFsm:
    """An NFA or DFA in flat CSR form.  All arrays are numpy, little-endian."""
    nstates: int
    start: int
    hasstart: bool
    is_end: np.ndarray            # u8 [nstates]
    group_off: np.ndarray         # u64 [nstates+1]
    group_symbols: np.ndarray     # u64 [ngroups, 4]
    group_to: np.ndarray          # u32 [ngroups]
    eps_off: np.ndarray           # u64 [nstates+1]
    eps_to: np.ndarray            # u32 [neps]
    endid_off: np.ndarray         # u64 [nstates+1]
    endids: np.ndarray            # u32 [nids]
    eager_off: np.ndarray | None = None   # u64 [nstates+1]: eager-output ids (None: the fsm has none)
    eager_ids: np.ndarray | None = None   # u32, sorted unique per state
    _keep: list = field(default_factory=list, repr=False, compare=False)

    def __post_init__(self):
        n = int(self.nstates)
        self.is_end = _arr(self.is_end, np.uint8)
        self.group_off = _arr(self.group_off, np.uint64)
        self.group_symbols = _arr(self.group_symbols, np.uint64).reshape(-1, 4)
        self.group_to = _arr(self.group_to, np.uint32)
        zeros = np.zeros(n + 1, dtype=np.uint64)
        self.eps_off = zeros.copy() if self.eps_off is None else _arr(self.eps_off, np.uint64)
        self.eps_to = np.zeros(0, np.uint32) if self.eps_to is None else _arr(self.eps_to, np.uint32)
        self.endid_off = zeros.copy() if self.endid_off is None else _arr(self.endid_off, np.uint64)
        self.endids = np.zeros(0, np.uint32) if self.endids is None else _arr(self.endids, np.uint32)
        if self.eager_off is not None:
            self.eager_off = _arr(self.eager_off, np.uint64)
            self.eager_ids = _arr(self.eager_ids if self.eager_ids is not None else [], np.uint32)
            assert self.eager_off.shape == (n + 1,) and int(self.eager_off[-1]) == self.eager_ids.shape[0]
            if self.eager_ids.size == 0:
                self.eager_off = self.eager_ids = None
        assert self.is_end.shape == (n,)
        assert self.group_off.shape == (n + 1,) and self.eps_off.shape == (n + 1,) and self.endid_off.shape == (n + 1,)
        assert int(self.group_off[-1]) == self.group_to.shape[0] == self.group_symbols.shape[0]

    # -- C view -------------------------------------------------------------------------
    def as_c(self) -> CDesc:
        """A ``struct fsm_b200_desc`` pointing into this object's arrays (keep ``self`` alive)."""
        def ptr(a):
            if a.size == 0:
                a = np.zeros(4, dtype=a.dtype)      # never hand NULL for an empty array
                self._keep.append(a)
            return a.ctypes.data
        ext = CDescExt() if self.eager_off is not None else None
        d = ext.base if ext is not None else CDesc()     # ext.base shares ext's memory and keeps it alive
        d.nstates = int(self.nstates); d.start = int(self.start)
        d.hasstart = 1 if self.hasstart else 0; d.reserved = DESC_EAGER if ext is not None else 0
        if ext is not None:
            ext.eager_off = ptr(self.eager_off); ext.eager_ids = ptr(self.eager_ids)
        d.is_end = ptr(self.is_end); d.group_off = ptr(self.group_off)
        d.group_symbols = ptr(self.group_symbols); d.group_to = ptr(self.group_to)
        d.eps_off = ptr(self.eps_off); d.eps_to = ptr(self.eps_to)
        d.endid_off = ptr(self.endid_off); d.endids = ptr(self.endids)
        return d

    @staticmethod
    def from_c(d: CDesc) -> "FlatFsm":
        """Deep copy out of a C description (so the C side can be freed)."""
        n = int(d.nstates)

        def take(p, count, dtype):
            if not p or count == 0:
                return np.zeros(0, dtype=dtype)
            buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(p)
            return np.frombuffer(buf, dtype=dtype).copy()
        group_off = take(d.group_off, n + 1, np.uint64) if n >= 0 else np.zeros(1, np.uint64)
        ng = int(group_off[-1]) if group_off.size else 0
        eps_off = take(d.eps_off, n + 1, np.uint64) if d.eps_off else np.zeros(n + 1, np.uint64)
        ne = int(eps_off[-1]) if eps_off.size else 0
        endid_off = take(d.endid_off, n + 1, np.uint64) if d.endid_off else np.zeros(n + 1, np.uint64)
        ni = int(endid_off[-1]) if endid_off.size else 0
        if group_off.size == 0:
            group_off = np.zeros(n + 1, np.uint64)
        return FlatFsm(
            nstates=n, start=int(d.start), hasstart=bool(d.hasstart),
            is_end=take(d.is_end, n, np.uint8),
            group_off=group_off,
            group_symbols=take(d.group_symbols, 4 * ng, np.uint64).reshape(-1, 4),
            group_to=take(d.group_to, ng, np.uint32),
            eps_off=eps_off, eps_to=take(d.eps_to, ne, np.uint32),
            endid_off=endid_off, endids=take(d.endids, ni, np.uint32))

    # -- helpers ------------------------------------------------------------------------
    def eager_of(self, state: int) -> np.ndarray:
        if self.eager_off is None:
            return np.zeros(0, np.uint32)
        return self.eager_ids[int(self.eager_off[state]):int(self.eager_off[state + 1])]

    def endids_of(self, state: int) -> np.ndarray:
        return self.endids[int(self.endid_off[state]):int(self.endid_off[state + 1])]

    def dense_table(self) -> np.ndarray:
        """[nstates, 256] uint32, 0xFFFFFFFF = no edge; first matching group wins
        (edge_set_find, reference src/adt/edgeset.c:394-418).  Host-side helper for
        tests and table inspection; not an execution path."""
        t = np.full((self.nstates, 256), 0xFFFFFFFF, dtype=np.uint32)
        bits = np.arange(256)
        for s in range(self.nstates):
            for g in range(int(self.group_off[s + 1]) - 1, int(self.group_off[s]) - 1, -1):
                sym = self.group_symbols[g]
                mask = ((sym[bits >> 6] >> (bits & 63).astype(np.uint64)) & np.uint64(1)).astype(bool)
                t[s, mask] = self.group_to[g]
        return t

    def save(self, path) -> None:
        np.savez_compressed(path, nstates=self.nstates, start=self.start, hasstart=int(self.hasstart),
                            is_end=self.is_end, group_off=self.group_off, group_symbols=self.group_symbols,
                            group_to=self.group_to, eps_off=self.eps_off, eps_to=self.eps_to,
                            endid_off=self.endid_off, endids=self.endids)

    @staticmethod
    def load(path) -> "FlatFsm":
        z = np.load(path)
        return FlatFsm(nstates=int(z["nstates"]), start=int(z["start"]), hasstart=bool(int(z["hasstart"])),
                       is_end=z["is_end"], group_off=z["group_off"], group_symbols=z["group_symbols"],
                       group_to=z["group_to"], eps_off=z["eps_off"], eps_to=z["eps_to"],
                       endid_off=z["endid_off"], endids=z["endids"])

    @staticmethod
    def from_edges(nstates, start, ends, edges, eps=(), endids=None, eager=None) -> "FlatFsm":
        """Build from explicit (src, symbol|iterable of symbols, dst) edges; groups are
        formed per (src, dst) and kept sorted by dst like the reference's edge_set."""
        groups = [dict() for _ in range(nstates)]
        for (s, sym, t) in edges:
            syms = [sym] if isinstance(sym, int) else list(sym)
            m = groups[s].setdefault(int(t), [0, 0, 0, 0])
            for c in syms:
                m[c >> 6] |= 1 << (c & 63)
        goff, gsym, gto = [0], [], []
        for s in range(nstates):
            for t in sorted(groups[s]):
                gsym.append(groups[s][t]); gto.append(t)
            goff.append(len(gto))
        eoff, eto = [0], []
        by_src = [[] for _ in range(nstates)]
        for (s, t) in eps:
            by_src[s].append(int(t))
        for s in range(nstates):
            eto.extend(sorted(set(by_src[s]))); eoff.append(len(eto))
        is_end = np.zeros(nstates, np.uint8)
        for e in ends:
            is_end[e] = 1
        ioff, ids = [0], []
        for s in range(nstates):
            if endids and s in endids and is_end[s]:
                ids.extend(sorted(set(endids[s])))
            ioff.append(len(ids))
        return FlatFsm(nstates=nstates, start=0 if start is None else start, hasstart=start is not None,
                       is_end=is_end, group_off=np.array(goff, np.uint64),
                       group_symbols=np.array(gsym, dtype=np.uint64).reshape(-1, 4),
                       group_to=np.array(gto, np.uint32), eps_off=np.array(eoff, np.uint64),
                       eps_to=np.array(eto, np.uint32), endid_off=np.array(ioff, np.uint64),
                       endids=np.array(ids, np.uint32), **FlatFsm._eager_csr(nstates, eager))

    @staticmethod
    def _eager_csr(nstates, eager) -> dict:
        """{state: ids} -> eager_off / eager_ids keyword arguments (empty when there are none)."""
        if not eager:
            return {}
        off, ids = [0], []
        for s in range(nstates):
            ids.extend(sorted(set(int(x) for x in eager.get(s, ()))))
            off.append(len(ids))
        return {"eager_off": np.array(off, np.uint64), "eager_ids": np.array(ids, np.uint32)}

    def with_eager(self, eager) -> "FlatFsm":
        """A copy carrying the eager-output sets {state: ids}."""
        import dataclasses
        kw = FlatFsm._eager_csr(self.nstates, eager) or {"eager_off": None, "eager_ids": None}
        return dataclasses.replace(self, _keep=[], **kw)
