"""CPU: the address arithmetic of the per-lane replicated tables (libfsm_b200/csrc/k1b_rep.cuh and
k1b_rep_tma.cuh), restated in Python from the comments of those files and checked exhaustively:
every (state, byte, lane) gets its own byte, the bank of an entry depends on the lane alone, the address
the kernel builds with one prmt.b32 per byte equals the address the table was written at, and the
chunking covers the stream exactly once.  (The kernels themselves are checked on the GPU:
tests/test_gpu_stream_rep.py.)"""
import numpy as np
import pytest


def prmt(a, b, sel):
    """prmt.b32 in its default mode: result byte k = byte (sel nibble k & 7) of {b, a} (a = bytes 0..3), or that
    byte's sign replicated when bit 3 of the nibble is set."""
    src = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
    out = 0
    for k in range(4):
        n = (sel >> (4 * k)) & 0xF
        v = src[n & 7]
        if n & 8:
            v = 0xFF if v & 0x80 else 0x00
        out |= v << (8 * k)
    return out


SEL = (0xCC40, 0xCC51, 0xCC62, 0xCC73)


def gap_build_addr(s, j, lane):                 # k1b_rep.cuh: word j of row s of lane l
    return (s << 14) | (j << 8) | (lane << 2)


def gap_step_addr(st, w, i, lane):
    h = (w >> 2) & 0x3F3F3F3F
    q = (w & 0x03030303) | ((lane << 2) * 0x01010101)
    return (st << 14) + prmt(q, h, SEL[i])


def compact_build_addr(s, j, lane):             # k1b_rep_tma.cuh
    return (s << 13) | ((j >> 1) << 8) | ((j & 1) << 7) | (lane << 2)


def compact_step_addr(st, w, i, lane):
    h = (w >> 3) & 0x1F1F1F1F
    q = ((w & 0x03030303) | ((lane << 2) * 0x01010101)) | ((w << 5) & 0x80808080)
    return (st << 13) + prmt(q, h, SEL[i])


@pytest.mark.parametrize("build,step,row_bytes", [(gap_build_addr, gap_step_addr, 1 << 14), (compact_build_addr, compact_step_addr, 1 << 13)])
def test_lookup_address_is_where_the_entry_was_written(build, step, row_bytes):
    rng = np.random.default_rng(1)
    seen = set()
    for st in (0, 1, 7, 11):
        for lane in range(32):
            for b in range(256):
                # the byte sits at position i of a word whose other bytes are random
                i = b & 3
                other = int(rng.integers(0, 1 << 32))
                w = (other & ~(0xFF << (8 * i))) | (b << (8 * i))
                got = step(st, w, i, lane)
                want = build(st, b >> 2, lane) + (b & 3)
                assert got == want, (st, lane, b)
                assert (got >> 2) & 31 == lane, "bank = lane"
                assert st * row_bytes <= got < (st + 1) * row_bytes
                assert got not in seen
                seen.add(got)


def test_twelve_rows_fit_the_shared_memory_budget():
    maps = 32 * 32 * 8 + 32 * 16 * 6 + 256          # REP_MAPS_BYTES
    assert maps + 16384 + 12 * (1 << 14) <= 232448   # 227 KiB opt-in limit of sm_100
    head = maps + 32 * 4 * 8                         # + full barriers (TMA form)
    ring_off = (head + 8192 + 9 * (1 << 13) + 1023) & ~1023
    assert ring_off + 4 * 32 * 1024 <= 232448        # UTF-8 validator (9 rows): four stages


@pytest.mark.parametrize("length,mis,sms", [(1, 0, 148), (63, 5, 148), (576, 0, 148), (300007, 31, 148), (1 << 31, 0, 148),
                                            ((1 << 31) + 12345, 17, 148), (10 ** 6, 3, 4)])
def test_chunks_cover_the_stream_once(length, mis, sms, T=8, W=64):
    """stream_map_rep's chunking (k1b_stream.cu): chunk 0 starts at offset 0, chunk c > 0 at c * C - mis (a sector
    boundary of the address), one chunk per lane of one wave at most."""
    lanes = sms * 1024
    C = ((length + mis + lanes - 1) // lanes + 31) & ~31
    cmin = (max(256, T * W) + 31) & ~31
    C = max(C, cmin)
    nchunks = (length + mis + C - 1) // C
    assert nchunks <= lanes and C % 32 == 0
    pos = 0
    for c in (range(nchunks) if nchunks < 5000 else list(range(3)) + list(range(nchunks - 3, nchunks))):
        beg = 0 if c == 0 else c * C - mis
        end = min(length, (c + 1) * C - mis)
        assert beg < end <= length
        if c > 0:
            assert (beg + mis) % 32 == 0
        if nchunks < 5000:
            assert beg == pos
            pos = end
    assert min(length, nchunks * C - mis) == length
