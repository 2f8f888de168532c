"""NVLink peer-memory plumbing for the fused scan + gather (one process per GPU).

Each rank allocates its `gathered` buffers with plain cudaMalloc (CUDA IPC maps whole
allocations), exports the IPC handles, exchanges them with one torch.distributed all-gather
and maps every peer's buffers.  `fsm_b200_exec_batch_dev_gather` then lets the scanning lanes
store each result record into every peer's buffer directly (P2P stores over NVLink/NVSwitch):
no collective kernel in the data path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._native import lib, check
from .desc import RESULT_DTYPE


class GatherRing:
    """`nbuf` gathered buffers of world*n result records on every rank, peer-mapped."""

    def __init__(self, n: int, world: int, rank: int, device: int, nbuf: int = 2, elem_bytes: int = 16):
        import torch
        import torch.distributed as dist
        self.n, self.world, self.rank, self.device, self.nbuf = n, world, rank, device, nbuf
        self.elem = elem_bytes                  # 16: full records; 4: compact match ids
        self.bytes = world * n * elem_bytes
        self.local = []
        for _ in range(nbuf):
            p = C.c_void_p()
            check(lib.fsm_b200_dev_alloc(device, self.bytes, C.byref(p)), "dev_alloc")
            self.local.append(p.value)
        handles = np.zeros((nbuf, 64), dtype=np.uint8)
        for b in range(nbuf):
            check(lib.fsm_b200_ipc_export(self.local[b], handles[b].ctypes.data), "ipc_export")
        mine = torch.from_numpy(handles.reshape(-1)).cuda(device)
        allh = torch.empty(world * mine.numel(), dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(allh, mine)
        allh = allh.cpu().numpy().reshape(world, nbuf, 64)
        self.peer = [[None] * world for _ in range(nbuf)]       # peer[b][r] = base of rank r's buffer b
        self._opened = []
        for b in range(nbuf):
            for r in range(world):
                if r == rank:
                    self.peer[b][r] = self.local[b]
                else:
                    p = C.c_void_p()
                    h = np.ascontiguousarray(allh[r, b])
                    check(lib.fsm_b200_ipc_open(device, h.ctypes.data, C.byref(p)), "ipc_open")
                    self.peer[b][r] = p.value
                    self._opened.append(p.value)
        self.slot = rank * n * elem_bytes

    def local_slot_ptr(self, b: int) -> int:
        """Where this rank's own records live inside its own gathered buffer b."""
        return self.local[b] + self.slot

    def peer_slot_ptrs(self, b: int):
        """ctypes array of the other ranks' gathered buffers b, at this rank's slot."""
        ptrs = [self.peer[b][r] + self.slot for r in range(self.world) if r != self.rank]
        return (C.c_void_p * len(ptrs))(*ptrs), len(ptrs)

    def read(self, b: int) -> np.ndarray:
        out = np.empty(self.world * self.n, dtype=RESULT_DTYPE if self.elem == 16 else np.uint32)
        check(lib.fsm_b200_dev_read(self.device, out.ctypes.data, self.local[b], self.bytes), "dev_read")
        return out

    def close(self) -> None:
        for p in self._opened:
            lib.fsm_b200_ipc_close(self.device, p)
        self._opened = []
        for p in self.local:
            lib.fsm_b200_dev_free(self.device, p)
        self.local = []
