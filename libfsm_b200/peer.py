"""NVLink peer-memory plumbing for the fused scan + gather (one process per GPU).

Each rank allocates its `gathered` buffers with plain cudaMalloc (CUDA IPC maps whole
allocations), exports the IPC handles, exchanges them with one torch.distributed all-gather
and maps every peer's buffers.  `fsm_b200_exec_batch_dev_gather` then lets the scanning lanes
store each result record (or 4-byte match id) into every peer's buffer directly (P2P stores
over NVLink/NVSwitch), and lets the kernel's last CTA publish a completion flag into every
peer's memory: no collective kernel anywhere in the data path.

Layout of one gathered buffer (per rank, per buffer index b):
    [ world * n * elem_bytes  records / ids ][ pad to 256 ][ world x u32 flags ][ u32 CTA counter ]
flag[r] on rank q == v  <=>  rank r's kernel that was launched with sig_value v has completed
and all of its stores into q's buffer are visible."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._native import lib, check
from .desc import RESULT_DTYPE


class GatherRing:
    """`nbuf` gathered buffers of world*n result records (or match ids) on every rank, peer-mapped."""

    def __init__(self, n: int, world: int, rank: int, device: int, nbuf: int = 2, elem_bytes: int = 16):
        import torch
        import torch.distributed as dist
        self.n, self.world, self.rank, self.device, self.nbuf = n, world, rank, device, nbuf
        self.elem = elem_bytes                  # 16: full records; 4: compact match ids
        self.bytes = world * n * elem_bytes
        self.flag_off = (self.bytes + 255) & ~255
        self.counter_off = self.flag_off + 256
        self.alloc_bytes = self.counter_off + 256
        self.local = []
        for _ in range(nbuf):
            p = C.c_void_p()
            check(lib.fsm_b200_dev_alloc(device, self.alloc_bytes, C.byref(p)), "dev_alloc")
            check(lib.fsm_b200_dev_zero(device, p.value + self.flag_off, 512), "dev_zero")
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
        self.others = [r for r in range(world) if r != rank]

    def local_slot_ptr(self, b: int) -> int:
        """Where this rank's own records live inside its own gathered buffer b."""
        return self.local[b] + self.slot

    def peer_slot_ptrs(self, b: int):
        """ctypes array of the other ranks' gathered buffers b, at this rank's slot."""
        ptrs = [self.peer[b][r] + self.slot for r in self.others]
        return (C.c_void_p * max(len(ptrs), 1))(*ptrs), len(ptrs)

    def signal_args(self, b: int):
        """(counter pointer, ctypes array of flag pointers): this rank's flag word inside every
        peer's buffer b (same order as peer_slot_ptrs), then inside its own."""
        flags = [self.peer[b][r] + self.flag_off + 4 * self.rank for r in self.others]
        flags.append(self.local[b] + self.flag_off + 4 * self.rank)
        return self.local[b] + self.counter_off, (C.c_void_p * len(flags))(*flags)

    def read(self, b: int) -> np.ndarray:
        out = np.empty(self.world * self.n, dtype=RESULT_DTYPE if self.elem == 16 else np.uint32)
        check(lib.fsm_b200_dev_read(self.device, out.ctypes.data, self.local[b], self.bytes), "dev_read")
        return out

    def read_flags(self, b: int) -> np.ndarray:
        """flag[r] for every source rank r, as seen in this rank's buffer b."""
        out = np.empty(self.world, dtype=np.uint32)
        check(lib.fsm_b200_dev_read(self.device, out.ctypes.data, self.local[b] + self.flag_off, 4 * self.world), "dev_read")
        return out

    def wait_flags(self, b: int, value: int) -> None:
        """Enqueue (current torch stream) the consumer's poll kernel: returns on the device once every
        rank's completion flag in this rank's buffer b has reached `value`.  A poll that gives up (a
        peer died) sets the word read by ``timed_out``."""
        import torch
        check(lib.fsm_b200_wait_flags_dev(self.device, self.local[b] + self.flag_off, self.world, int(value) & 0xFFFFFFFF,
                                          self.local[b] + self.counter_off + 64, int(torch.cuda.current_stream().cuda_stream)),
              "wait_flags_dev")

    def timed_out(self) -> bool:
        for b in range(self.nbuf):
            out = np.empty(1, dtype=np.uint32)
            check(lib.fsm_b200_dev_read(self.device, out.ctypes.data, self.local[b] + self.counter_off + 64, 4), "dev_read")
            if out[0]:
                return True
        return False

    def close(self) -> None:
        for p in self._opened:
            lib.fsm_b200_ipc_close(self.device, p)
        self._opened = []
        for p in self.local:
            lib.fsm_b200_dev_free(self.device, p)
        self.local = []
