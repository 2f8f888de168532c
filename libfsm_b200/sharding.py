"""Multi-GPU plumbing for the hot path (one process per GPU, torch.distributed).

Batches (configs 2/3): inputs are independent, so the index space is cut into contiguous
ranges balanced by BYTES; every rank scans its range with a full copy of the table and the
fixed-size result records are exchanged with ONE all-gather.  No data-path collective.

One long stream (config 4): DFA execution is a monoid.  Every rank scans its byte range and
produces, for every possible entry state, (exit state, first dead offset, state it died in);
ONE all-gather of these [ntable] records, then each rank composes them in rank order
(`compose_stream_maps`) -- exact for any DFA, no reliance on self-synchronisation.
"""
from __future__ import annotations

import numpy as np

NO_DEAD = np.uint64(0xFFFFFFFFFFFFFFFF)


def shard_ranges_by_bytes(offsets: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous input ranges [lo, hi) per rank with ~equal byte counts.  offsets: u64 [n+1]."""
    n = len(offsets) - 1
    total = int(offsets[-1] - offsets[0])
    cuts = [0]
    for r in range(1, world):
        target = int(offsets[0]) + total * r // world
        k = int(np.searchsorted(offsets, target, side="left"))
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_range_fixed(n: int, world: int, rank: int) -> tuple[int, int]:
    return n * rank // world, n * (rank + 1) // world


def byte_ranges(nbytes: int, world: int, align: int = 16) -> list[tuple[int, int]]:
    """Byte ranges of one long stream, cut at `align`-byte boundaries."""
    cuts = [0]
    for r in range(1, world):
        c = (nbytes * r // world) // align * align
        cuts.append(max(c, cuts[-1]))
    cuts.append(nbytes)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def compose_stream_maps(start: int, dead_row: int | None, shard_lens, map_state, map_dead, map_dead_state):
    """Fold per-shard maps in order.  map_*[r][s] as returned by Dfa.exec_stream_map for shard r.
    Returns (final_state, consumed, died): `consumed` is the global offset of the first byte
    without an edge when died, else the total length."""
    state, base = int(start), 0
    for r, ln in enumerate(shard_lens):
        d = np.uint64(map_dead[r][state])
        if d != NO_DEAD:
            return int(map_dead_state[r][state]), base + int(d), True
        state = int(map_state[r][state])
        if dead_row is not None and state == dead_row:        # defensive: never expected
            raise AssertionError("dead state without a recorded offset")
        base += int(ln)
    return state, base, False
