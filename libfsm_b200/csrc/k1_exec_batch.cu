/*
 * k1_exec_batch.cu -- K1: batched DFA execution, one reference fsm_exec call per input.
 *
 * Replaces the per-byte loop of fsm_exec (src/libfsm/exec.c:132-151) and its
 * edge_set_transition/edge_set_find group scan (src/adt/edgeset.c:565-579, 394-418)
 * by a dense-table walk: state = T[state][byte].  One lane walks one input (inputs are
 * independent, so this does no redundant work and needs no speculation); the serial
 * dependence is hidden by the thousands of other lanes resident on the SM.
 *
 * Variants (DESIGN.md "K1 variants"; chosen by fsm_b200_set_exec_variant / auto):
 *   LANE   k1_lane_kernel   any layout (ragged offsets, unaligned).  Table TMA-bulk-staged
 *                           to shared memory (or read from L2 when too big); input read
 *                           straight from HBM as 256-bit loads per lane.
 *   TILE   k1_tile_kernel   fixed-stride, 16B-aligned batches.  Per warp a ring of
 *                           32-row x CH-byte input tiles is fetched by 2-D TMA
 *                           (cp.async.bulk.tensor, hardware swizzle so the lane-strided
 *                           16-byte reads are bank-conflict free) and completed on mbarriers.
 *
 * Integer/byte workload: HBM- and shared-memory-lookup-bound; tensor cores do not apply.
 */
#include <cstring>
#include <cuda.h>
#include <mutex>

#include "common.h"
#include "k1_exec_batch.h"
#include "k1_device.cuh"

using namespace fsmb200;

namespace {

/* ------------------------------------------------------------------ device helpers -- */

/* 2-D TMA tensor tile global -> shared (SASS: UTMALDG). */
__device__ __forceinline__ void
tma_tile_g2s(uint32_t dst, const CUtensorMap *map, uint32_t c0, uint32_t c1, uint32_t bar)
{
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes"
	    " [%0], [%1, {%2, %3}], [%4];"
	    :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

/* Table lookups.  Indexing the extern shared array directly lets ptxas emit, per input
 * byte, exactly PRMT (byte extract, off the dependent chain) + IMAD (state*pitch + byte)
 * + LDS.U8 [R + UR] -- the dependent chain is IMAD -> LDS. */
template <typename E> struct TableSmem {
	const E *tbl;        /* shared memory */
	uint32_t pitch;      /* row pitch in ELEMENTS */
	__device__ __forceinline__ uint32_t step(uint32_t st, uint32_t b) const {
		return (uint32_t) tbl[st * pitch + b];
	}
};

template <typename E> struct TableGmem {
	const E *tbl;        /* global memory (L2-resident) */
	uint32_t pitch;      /* row pitch in ELEMENTS */
	__device__ __forceinline__ uint32_t step(uint32_t st, uint32_t b) const {
		return (uint32_t) __ldg(tbl + (size_t) st * pitch + b);
	}
};

#define STEP4(T, st, w)                                    \
	do {                                                   \
		st = (T).step(st, __byte_perm((w), 0u, 0x4440u));  \
		st = (T).step(st, __byte_perm((w), 0u, 0x4441u));  \
		st = (T).step(st, __byte_perm((w), 0u, 0x4442u));  \
		st = (T).step(st, __byte_perm((w), 0u, 0x4443u));  \
	} while (0)

/* One 16-byte record per input; with the fused gather the same record is also stored into
 * every peer GPU's gathered buffer (P2P stores over NVLink, posted: they overlap the scan). */
__device__ __forceinline__ void
store_result(const K1Args &a, uint64_t i, int32_t ret, uint32_t end, uint64_t consumed)
{
	uint4 v;
	v.x = (uint32_t) ret;
	v.y = end;
	v.z = (uint32_t) consumed;
	v.w = (uint32_t) (consumed >> 32);
	*reinterpret_cast<uint4 *>(a.out + i) = v;
	if (a.peer_compact) {
		const uint32_t id = (ret == 1 ? 0x80000000u : 0u) | end;      /* the match id: 4 B on the wire */
		for (uint32_t r = 0; r < a.npeers; r++) {
			reinterpret_cast<uint32_t *>(a.peer_out[r])[i] = id;
		}
	} else {
		for (uint32_t r = 0; r < a.npeers; r++) {
			*reinterpret_cast<uint4 *>(a.peer_out[r] + i) = v;
		}
	}
}

/* Fused-gather completion signal.  Every CTA: all its P2P stores issued -> __threadfence_system
 * -> count itself done.  The last CTA resets the counter and publishes sig_value into the flag
 * word of every peer (and its own): a consumer that sees flag == value from rank r knows all of
 * r's records for that step have landed in its gathered buffer.  No collective kernel needed. */
__device__ __forceinline__ void
signal_done(const K1Args &a)
{
	if (a.sig_counter == nullptr) return;
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence_system();
		const uint32_t old = atomicAdd(a.sig_counter, 1u);
		if (old == gridDim.x - 1) {
			atomicExch(a.sig_counter, 0u);
			__threadfence_system();
			for (uint32_t r = 0; r <= a.npeers; r++) {
				*reinterpret_cast<volatile uint32_t *>(a.sig_flags[r]) = a.sig_value;
			}
			__threadfence_system();
		}
	}
}

/* ------------------------------------------------------------------ LANE variant ---- */

template <typename E, bool SMEM, bool HAS_DEAD, bool CLS>
__global__ void __launch_bounds__(1024, 1)
k1_lane_kernel(const K1Args a)
{
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint64_t blob_bar;

	const uint8_t *is_end;
	const uint8_t *cls = nullptr;      /* CLS: byte -> class LUT; rows are indexed by class */
	TableSmem<E> ts;
	TableGmem<E> tg;
	if (SMEM) {
		stage_blob(smem, a.blob, a.blob_bytes, &blob_bar);
		ts.tbl = reinterpret_cast<const E *>(smem);
		ts.pitch = a.pitch / (uint32_t) sizeof(E);
		is_end = smem + a.is_end_off;
		if (CLS) cls = smem + a.cls_off;
	} else {
		tg.tbl = reinterpret_cast<const E *>(a.blob);
		tg.pitch = a.pitch / (uint32_t) sizeof(E);
		is_end = a.blob + a.is_end_off;
		if (CLS) cls = a.blob + a.cls_off;       /* 256 B, L1-resident */
	}
#define COL(b) (CLS ? (uint32_t) (SMEM ? cls[(b)] : __ldg(cls + (b))) : (uint32_t) (b))
#define TSTEP(st, b) (SMEM ? ts.step(st, COL(b)) : tg.step(st, COL(b)))
#define WSTEP4(st, w)                                          \
	do {                                                       \
		st = TSTEP(st, __byte_perm((w), 0u, 0x4440u));         \
		st = TSTEP(st, __byte_perm((w), 0u, 0x4441u));         \
		st = TSTEP(st, __byte_perm((w), 0u, 0x4442u));         \
		st = TSTEP(st, __byte_perm((w), 0u, 0x4443u));         \
	} while (0)

	const uint64_t nthreads = (uint64_t) gridDim.x * blockDim.x;
	const uint64_t n_inputs = a.n_dev != nullptr ? (uint64_t) *a.n_dev : a.n;
	for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_inputs; i += nthreads) {
		uint64_t beg, len;
		if (a.offsets != nullptr) {
			beg = a.offsets[i];
			len = (a.ends != nullptr ? a.ends[i] : a.offsets[i + 1]) - beg;
		} else {
			beg = i * a.stride;
			len = a.len;
		}
		const uint8_t *p = a.base + beg;
		uint32_t st = a.entry != nullptr ? a.entry[i] : a.start;
		uint64_t pos = 0;
		bool died = false;

		/* head: single bytes up to the first 32-byte boundary */
		uint64_t head = (uint64_t) ((32u - (uint32_t) (reinterpret_cast<uintptr_t>(p) & 31u)) & 31u);
		if (head > len) head = len;
		for (; pos < head; pos++) {
			const uint32_t nx = TSTEP(st, (uint32_t) __ldg(p + pos));
			if (HAS_DEAD && nx == a.dead) { died = true; break; }
			st = nx;
		}

		if (!died) {
			/* body: one full 32-byte sector per 256-bit load, next sector prefetched */
			uint64_t nchunk = (len - pos) >> 5;
			uint32_t cur[8], nxt[8];
			if (nchunk > 0) ld256(p + pos, cur);
			for (uint64_t c = 0; c < nchunk; c++) {
				if (c + 1 < nchunk) {
					ld256(p + pos + 32, nxt);
				} else {
#pragma unroll
					for (int k = 0; k < 8; k++) nxt[k] = 0;
				}
				const uint32_t entry = st;
#pragma unroll
				for (int k = 0; k < 8; k++) WSTEP4(st, cur[k]);
				if (HAS_DEAD && st == a.dead) {
					/* a byte of this sector had no edge: re-walk it to find which */
					st = entry;
					for (int k = 0; k < 32; k++) {
						const uint32_t nx = TSTEP(st, (uint32_t) __ldg(p + pos + k));
						if (nx == a.dead) { died = true; pos += (uint64_t) k; break; }
						st = nx;
					}
					break;
				}
				pos += 32;
				/* absorbing state (all 256 edges loop back): the rest of the input cannot change
				 * the record -- stop reading it */
				if (a.absorb != nullptr && __ldg(a.absorb + st)) { pos = len; break; }
#pragma unroll
				for (int k = 0; k < 8; k++) cur[k] = nxt[k];
			}
		}
		if (!died) {
			for (; pos < len; pos++) {
				const uint32_t nx = TSTEP(st, (uint32_t) __ldg(p + pos));
				if (HAS_DEAD && nx == a.dead) { died = true; break; }
				st = nx;
			}
		}
		const int32_t ret = (!died && is_end[st]) ? 1 : 0;
		store_result(a, i, ret, st, pos);
	}
	signal_done(a);
#undef TSTEP
#undef WSTEP4
#undef COL
}

/* First byte of a sector (within `mask`) whose transition enters the dead row, and the state
 * it was taken from.  Returns (from_state, byte index). */
template <typename T, bool CLS, bool SMEM>
__device__ __noinline__ uint2
rewalk_sector(const T tab, const uint8_t *cls, uint32_t entry, uint32_t mask, uint32_t dead, const uint32_t (&w)[8])
{
	uint32_t st = entry, at = 0, from = entry;
	bool found = false;
#pragma unroll 1
	for (int k = 0; k < 8; k++) {
		const uint32_t word = w[k];
#pragma unroll
		for (int t = 0; t < 4; t++) {
			const uint32_t j = 4 * k + t;
			const uint32_t b = (word >> (8 * t)) & 0xFFu;
			const uint32_t col = CLS ? (uint32_t) (SMEM ? cls[b] : __ldg(cls + b)) : b;
			const uint32_t nx = tab.step(st, col);
			const bool in = ((mask >> j) & 1u) != 0u;
			const bool hit = in && !found && nx == dead;
			at = hit ? j : at;
			from = hit ? st : from;
			found = found || hit;
			st = (in && !found) ? nx : st;
		}
	}
	return make_uint2(from, at);
}

/* ------------------------------------------------------------------ RAGGED variant -- */

/*
 * Ragged batches (offsets array, arbitrary alignment, lines of 10..1000 bytes that may die
 * after a few bytes: BASELINE config 3).  Lane l of a warp walks lines l, l+32, l+64, ... of
 * the warp's contiguous range, ONE aligned 32-byte sector per loop iteration:
 *   - every load is a full aligned sector (256-bit), also for unaligned line starts/ends:
 *     bytes outside [lo, hi) of the sector are walked too but their result is discarded
 *     with a select (a lookup with any byte value is in bounds), so there is no byte-wise
 *     head/tail path and no dependent global load in the chain;
 *   - three sector buffers: A (being walked), B (next sector of the same line) and N (first
 *     sector of the lane's NEXT line, fetched when the current line starts), so a lane that
 *     finishes -- or dies in -- a line continues immediately with a prefetched sector;
 *   - a died line is re-walked from registers (same select scheme) to find the exact offset.
 * Sectors that would reach outside [base+offsets[0], base+offsets[n]) (only the first and the
 * last line of a batch can) are assembled from byte loads instead.
 */
template <typename E, bool SMEM, bool HAS_DEAD, bool CLS>
__global__ void __launch_bounds__(768, 1)      /* three sector buffers: 85 registers per lane */
k1_ragged_kernel(const K1Args a)
{
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint64_t blob_bar;

	const uint8_t *is_end;
	const uint8_t *cls = nullptr;
	TableSmem<E> ts;
	TableGmem<E> tg;
	if (SMEM) {
		stage_blob(smem, a.blob, a.blob_bytes, &blob_bar);
		ts.tbl = reinterpret_cast<const E *>(smem);
		ts.pitch = a.pitch / (uint32_t) sizeof(E);
		is_end = smem + a.is_end_off;
		if (CLS) cls = smem + a.cls_off;
	} else {
		tg.tbl = reinterpret_cast<const E *>(a.blob);
		tg.pitch = a.pitch / (uint32_t) sizeof(E);
		is_end = a.blob + a.is_end_off;
		if (CLS) cls = a.blob + a.cls_off;
	}
#define COL(b) (CLS ? (uint32_t) (SMEM ? cls[(b)] : __ldg(cls + (b))) : (uint32_t) (b))
#define TSTEP(st, b) (SMEM ? ts.step(st, COL(b)) : tg.step(st, COL(b)))

	const uint32_t lane = threadIdx.x & 31u;
	const uint64_t nwarps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
	const uint64_t gw = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	/* contiguous range of lines per warp, in multiples of 32 */
	const uint64_t per = (((a.n + nwarps - 1) / nwarps) + 31u) & ~31ull;
	const uint64_t wbeg = gw * per;
	const uint64_t wend = min(a.n, wbeg + per);
	const uintptr_t lo_ptr = reinterpret_cast<uintptr_t>(a.base) + a.offsets[0];
	const uintptr_t hi_ptr = reinterpret_cast<uintptr_t>(a.base) + a.offsets[a.n];

	auto load_sector = [&](uintptr_t saddr, uint32_t (&w)[8]) {
		if (saddr >= lo_ptr && saddr + 32 <= hi_ptr) {
			ld256(reinterpret_cast<const uint8_t *>(saddr), w);
		} else {                                     /* batch edge: assemble from byte loads */
#pragma unroll
			for (int k = 0; k < 8; k++) {
				uint32_t v = 0;
#pragma unroll
				for (int t = 0; t < 4; t++) {
					const uintptr_t p = saddr + 4 * k + t;
					if (p >= lo_ptr && p < hi_ptr) v |= (uint32_t) __ldg(reinterpret_cast<const uint8_t *>(p)) << (8 * t);
				}
				w[k] = v;
			}
		}
	};

	uint64_t i = wbeg + lane;
	bool have = i < wend;
	uintptr_t cur = 0, end = 0, nbeg = 0, nend = 0;
	uint32_t A[8], B[8], N[8];
#pragma unroll
	for (int k = 0; k < 8; k++) { A[k] = 0; B[k] = 0; N[k] = 0; }
	bool have_next = false;
	if (have) {
		cur = reinterpret_cast<uintptr_t>(a.base) + a.offsets[i];
		end = reinterpret_cast<uintptr_t>(a.base) + (a.ends != nullptr ? a.ends[i] : a.offsets[i + 1]);
		load_sector(cur & ~(uintptr_t) 31, A);
		have_next = i + 32 < wend;
		if (have_next) {
			nbeg = reinterpret_cast<uintptr_t>(a.base) + a.offsets[i + 32];
			nend = reinterpret_cast<uintptr_t>(a.base) + (a.ends != nullptr ? a.ends[i + 32] : a.offsets[i + 33]);
			load_sector(nbeg & ~(uintptr_t) 31, N);
		}
	}
	uint32_t st = a.entry != nullptr && have ? a.entry[i] : a.start;
	uintptr_t line_beg = cur;

	while (have) {
		const uintptr_t saddr = cur & ~(uintptr_t) 31;
		const uint32_t lo = (uint32_t) (cur - saddr);
		const uint32_t hi = (end - saddr) >= 32 ? 32u : (uint32_t) (end - saddr);
		const bool more = saddr + 32 < end;              /* the line continues in the next sector */
		if (more) load_sector(saddr + 32, B);

		/* walk the 32 bytes; keep the new state only for bytes inside [lo, hi) */
		const uint32_t mask = (hi > lo) ? ((0xFFFFFFFFu << lo) & (0xFFFFFFFFu >> (32u - hi))) : 0u;
		const uint32_t entry = st;
#pragma unroll
		for (int k = 0; k < 8; k++) {
#pragma unroll
			for (int t = 0; t < 4; t++) {
				const uint32_t nx = TSTEP(st, __byte_perm(A[k], 0u, 0x4440u + t));
				st = ((mask >> (4 * k + t)) & 1u) ? nx : st;
			}
		}
		bool died = false;
		uint32_t consumed_here = hi - lo;
		if (HAS_DEAD && st == a.dead) {
			/* re-walk from registers to find the first byte without an edge (out of line: keeps
			 * its temporaries out of the main loop's register budget) */
			died = true;
			const uint2 r = SMEM ? rewalk_sector<TableSmem<E>, CLS, SMEM>(ts, cls, entry, mask, a.dead, A)
			                     : rewalk_sector<TableGmem<E>, CLS, SMEM>(tg, cls, entry, mask, a.dead, A);
			st = r.x;
			consumed_here = r.y - lo;
		}
		/* A state whose 256 edges all loop back to itself keeps the walk where it is whatever
		 * follows (the accept state of an end-unanchored pattern): the rest of the line cannot
		 * change the record, so it is neither walked nor read. */
		bool absorbed = false;
		if (!died && more && a.absorb != nullptr && __ldg(a.absorb + st)) {
			absorbed = true;
			cur = end;
		} else {
			cur += consumed_here;
		}
		if (died || !more || absorbed) {
			/* line done */
			const int32_t ret = (!died && is_end[st]) ? 1 : 0;
			store_result(a, i, ret, st, (uint64_t) (cur - line_beg));
			i += 32;
			have = have_next;
			if (have) {
				cur = nbeg; end = nend; line_beg = cur;
				st = a.entry != nullptr ? a.entry[i] : a.start;
#pragma unroll
				for (int k = 0; k < 8; k++) A[k] = N[k];
				have_next = i + 32 < wend;
				if (have_next) {
					nbeg = reinterpret_cast<uintptr_t>(a.base) + a.offsets[i + 32];
					nend = reinterpret_cast<uintptr_t>(a.base) + (a.ends != nullptr ? a.ends[i + 32] : a.offsets[i + 33]);
					load_sector(nbeg & ~(uintptr_t) 31, N);
				}
			}
		} else {
			cur = saddr + 32;
#pragma unroll
			for (int k = 0; k < 8; k++) A[k] = B[k];
		}
	}
#undef TSTEP
#undef COL
}

/* ------------------------------------------------------------------ K-STRIDE variant -- */

/*
 * Lane-per-input like LANE, but the dependent chain advances K bytes per lookup:
 *   idx = L0[b0] + L1[b1] (+ L2[b2] + L3[b3])     K independent 256-byte class LUT reads
 *   st  = stepK[st * pitch + idx]                  one dependent read per K bytes
 * Heads, tails and the re-walk of a sector that died use the single-byte class table.
 * 8-bit entries, everything in shared memory (a few KB to a few tens of KB).
 */
template <int K, bool HAS_DEAD>
__global__ void __launch_bounds__(1024, 1)
k1_kstride_kernel(const K1Args a)
{
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint64_t blob_bar;
	stage_blob(smem, a.kblob, a.kblob_bytes, &blob_bar);

	/* blob layout: [K LUTs][stepK rows][step1 rows][is_end] (dfa_compile.cu) */
	const uint8_t *L0 = smem, *L1 = smem + 256, *L2 = smem + 512, *L3 = smem + 768;
	const uint8_t *tk = smem + 256 * K;
	const uint8_t *t1 = smem + a.k1_off;
	const uint8_t *is_end = smem + a.kend_off;
	const uint32_t kp = a.kpitch, p1 = a.k1pitch;
#define STEP1(st, b) ((uint32_t) t1[(st) * p1 + L0[(b)]])

	const uint64_t nthreads = (uint64_t) gridDim.x * blockDim.x;
	const uint64_t n_inputs = a.n_dev != nullptr ? (uint64_t) *a.n_dev : a.n;
	for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_inputs; i += nthreads) {
		uint64_t beg, len;
		if (a.offsets != nullptr) {
			beg = a.offsets[i];
			len = (a.ends != nullptr ? a.ends[i] : a.offsets[i + 1]) - beg;
		} else {
			beg = i * a.stride;
			len = a.len;
		}
		const uint8_t *p = a.base + beg;
		uint32_t st = a.entry != nullptr ? a.entry[i] : a.start;
		uint64_t pos = 0;
		bool died = false;

		uint64_t head = (uint64_t) ((32u - (uint32_t) (reinterpret_cast<uintptr_t>(p) & 31u)) & 31u);
		if (head > len) head = len;
		for (; pos < head; pos++) {
			const uint32_t nx = STEP1(st, (uint32_t) __ldg(p + pos));
			if (HAS_DEAD && nx == a.dead) { died = true; break; }
			st = nx;
		}
		if (!died) {
			uint64_t nchunk = (len - pos) >> 5;
			uint32_t cur[8], nxt[8];
			if (nchunk > 0) ld256(p + pos, cur);
			for (uint64_t c = 0; c < nchunk; c++) {
				if (c + 1 < nchunk) {
					ld256(p + pos + 32, nxt);
				} else {
#pragma unroll
					for (int k = 0; k < 8; k++) nxt[k] = 0;
				}
				const uint32_t entry = st;
#pragma unroll
				for (int k = 0; k < 8; k++) {
					const uint32_t w = cur[k];
					const uint32_t c0 = L0[__byte_perm(w, 0u, 0x4440u)], c1 = L1[__byte_perm(w, 0u, 0x4441u)];
					if (K == 4) {
						const uint32_t c2 = L2[__byte_perm(w, 0u, 0x4442u)], c3 = L3[__byte_perm(w, 0u, 0x4443u)];
						st = tk[st * kp + (c0 + c1 + c2 + c3)];
					} else {
						const uint32_t d0 = L0[__byte_perm(w, 0u, 0x4442u)], d1 = L1[__byte_perm(w, 0u, 0x4443u)];
						st = tk[st * kp + (c0 + c1)];
						st = tk[st * kp + (d0 + d1)];
					}
				}
				if (HAS_DEAD && st == a.dead) {
					st = entry;
					for (int k = 0; k < 32; k++) {
						const uint32_t nx = STEP1(st, (uint32_t) __ldg(p + pos + k));
						if (nx == a.dead) { died = true; pos += (uint64_t) k; break; }
						st = nx;
					}
					break;
				}
				pos += 32;
				if (a.absorb != nullptr && __ldg(a.absorb + st)) { pos = len; break; }   /* absorbing: done */
#pragma unroll
				for (int k = 0; k < 8; k++) cur[k] = nxt[k];
			}
		}
		if (!died) {
			for (; pos < len; pos++) {
				const uint32_t nx = STEP1(st, (uint32_t) __ldg(p + pos));
				if (HAS_DEAD && nx == a.dead) { died = true; break; }
				st = nx;
			}
		}
		const int32_t ret = (!died && is_end[st]) ? 1 : 0;
		store_result(a, i, ret, st, pos);
	}
	signal_done(a);
#undef STEP1
}


/* ------------------------------------------------------------------ K-RANGE variant ---- */

/*
 * K-STRIDE with ALU byte classification (K = 4): the tuple index of a 4-byte word is computed in
 * registers, no LUT reads.  The byte class is a function of the 2-bit cell code
 * [b in R0] + 2 [b in R1] of two byte ranges (dfa_compile.cu: find_cell_ranges).  Per word, all four
 * bytes at once:
 *   l     = w & 0x7F7F7F7F                              7-bit values: the adds cannot carry across bytes
 *   in_k  = (l + add_lo_k) & ~(l + add_hi_k) & half_k   bit 7 of every byte: lo_k <= b <= hi_k
 *   i128  = dp4a(in_0, {1,4,16,64}) + dp4a(in_1, {2,8,32,128})        = 128 * tuple index
 *   st    = stepK_T[2 * i128 + st]                      table stored [tuple][state], 256 states per tuple
 * half_k keeps the bytes of the half of the byte space R_k lives in (bit 7 clear, or set); with
 * RNG == 1 both ranges lie below 0x80 and share it.  11 integer instructions + ONE shared-memory
 * wavefront per 4 bytes (the LUT form: 12 instructions, 5 wavefronts).
 * Three sector buffers in rotation (no register copies): loads run two sectors ahead of the walk.
 * Optionally (kr_prefetch != 0, off by default: it measured slower) the cache line that many bytes
 * ahead is requested into L2 as well.
 * blob: [256 B cell LUT][stepK_T 256 x 256][step1 rows][is_end]
 */
template <bool HAS_DEAD, int RNG>
__global__ void __launch_bounds__(1024, 1)
k1_krange_kernel(const K1Args a)
{
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint64_t blob_bar;
	stage_blob(smem, a.kblob, a.kblob_bytes, &blob_bar);

	const uint8_t *L0 = smem;
	const uint8_t *t1 = smem + a.k1_off;
	const uint8_t *is_end = smem + a.kend_off;
	const uint32_t p1 = a.k1pitch;
	const uint32_t pfd = a.kr_prefetch;
	const uint32_t kt_half = (smem_u32(smem) + 256u) >> 1;       /* dynamic shared memory is 1024-aligned */
#define STEP1(st, b) ((uint32_t) t1[(st) * p1 + L0[(b)]])

	const uint64_t nthreads = (uint64_t) gridDim.x * blockDim.x;
	const uint64_t n_inputs = a.n_dev != nullptr ? (uint64_t) *a.n_dev : a.n;
	for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_inputs; i += nthreads) {
		uint64_t beg, len;
		if (a.offsets != nullptr) {
			beg = a.offsets[i];
			len = (a.ends != nullptr ? a.ends[i] : a.offsets[i + 1]) - beg;
		} else {
			beg = i * a.stride;
			len = a.len;
		}
		const uint8_t *p = a.base + beg;
		uint32_t st = a.entry != nullptr ? a.entry[i] : a.start;
		uint64_t pos = 0;
		bool died = false;

		uint64_t head = (uint64_t) ((32u - (uint32_t) (reinterpret_cast<uintptr_t>(p) & 31u)) & 31u);
		if (head > len) head = len;
		for (; pos < head; pos++) {
			const uint32_t nx = STEP1(st, (uint32_t) __ldg(p + pos));
			if (HAS_DEAD && nx == a.dead) { died = true; break; }
			st = nx;
		}
		if (!died) {
			const uint64_t body_end = pos + ((len - pos) & ~31ull);      /* whole sectors: [pos, body_end) */
			uint32_t A[8], B[8], C[8];
			bool stop = false;
			/* walks one sector; stop = died, or absorbed */
#define KR_WALK(cur)                                                                                     \
			do {                                                                                         \
				const uint32_t entry = st;                                                               \
				_Pragma("unroll")                                                                        \
				for (int k = 0; k < 8; k++) {                                                            \
					const uint32_t w = (cur)[k];                                                         \
					const uint32_t l = w & 0x7F7F7F7Fu;                                                  \
					const uint32_t half0 = (w ^ a.kr_hxor[0]) & 0x80808080u;                             \
					const uint32_t half1 = RNG == 1 ? half0 : ((w ^ a.kr_hxor[1]) & 0x80808080u);        \
					const uint32_t in0 = (l + a.kr_add_lo[0]) & ~(l + a.kr_add_hi[0]) & half0;           \
					const uint32_t in1 = (l + a.kr_add_lo[1]) & ~(l + a.kr_add_hi[1]) & half1;           \
					/* 128 * tuple + half the shared-memory address of stepK_T: the lookup address */    \
					/* is st + 2 * acc, one IADD3, no separate base */                                   \
					const uint32_t acc = __dp4a(in1, 0x80200802u, __dp4a(in0, 0x40100401u, kt_half));    \
					asm("ld.shared.u8 %0, [%1];" : "=r"(st) : "r"(st + acc + acc));                      \
				}                                                                                        \
				if (HAS_DEAD && st == a.dead) {                                                          \
					/* a byte of this sector had no edge: re-walk it to find which */                    \
					st = entry;                                                                          \
					for (int k = 0; k < 32; k++) {                                                       \
						const uint32_t nx = STEP1(st, (uint32_t) __ldg(p + pos + k));                    \
						if (nx == a.dead) { died = true; pos += (uint64_t) k; break; }                   \
						st = nx;                                                                         \
					}                                                                                    \
					stop = true;                                                                         \
				} else {                                                                                 \
					pos += 32;                                                                           \
					if (a.absorb != nullptr && __ldg(a.absorb + st)) { pos = len; stop = true; }         \
				}                                                                                        \
			} while (0)
			/* first sector of a 128-byte line: ask L2 for the line `pfd` bytes ahead */
#define KR_FETCH(at, buf)                                                                                \
			do {                                                                                         \
				if ((at) < body_end) {                                                                   \
					ld256(p + (at), buf);                                                                \
					if (pfd != 0 && (reinterpret_cast<uintptr_t>(p + (at)) & 96u) == 0 && (at) + pfd < len) { \
						asm volatile("prefetch.global.L2 [%0];" :: "l"(p + (at) + pfd));                 \
					}                                                                                    \
				}                                                                                        \
			} while (0)
			/* three sector buffers in rotation: the loads run two sectors (64 B per lane, 9.7 MB per
			 * GPU) ahead of the walk -- with one sector ahead the bytes in flight, not DRAM or the issue
			 * rate, capped the kernel at 5.3 TB/s */
			KR_FETCH(pos, A);
			KR_FETCH(pos + 32, B);
			while (pos < body_end) {
				KR_FETCH(pos + 64, C);
				KR_WALK(A);
				if (stop || pos >= body_end) break;
				KR_FETCH(pos + 64, A);
				KR_WALK(B);
				if (stop || pos >= body_end) break;
				KR_FETCH(pos + 64, B);
				KR_WALK(C);
				if (stop) break;
			}
#undef KR_WALK
#undef KR_FETCH
		}
		if (!died) {
			for (; pos < len; pos++) {
				const uint32_t nx = STEP1(st, (uint32_t) __ldg(p + pos));
				if (HAS_DEAD && nx == a.dead) { died = true; break; }
				st = nx;
			}
		}
		const int32_t ret = (!died && is_end[st]) ? 1 : 0;
		store_result(a, i, ret, st, pos);
	}
	signal_done(a);
#undef STEP1
}

/* ------------------------------------------------------------------ TILE variant ---- */

/* Shared memory: [blob, padded to 1024][per warp: NSTAGE stages of 32 x CH bytes][mbarriers] */
template <typename E, bool HAS_DEAD, int CH, int NSTAGE>
__global__ void __launch_bounds__(1024, 1)
k1_tile_kernel(const K1Args a, const __grid_constant__ CUtensorMap tmap)
{
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint64_t blob_bar;
	constexpr uint32_t STAGE_BYTES = 32u * CH;
	constexpr int NVEC = CH / 16;

	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	const uint32_t nwarps = blockDim.x >> 5;
	uint8_t *stage_base = smem + a.tile_stage_off + (size_t) warp * NSTAGE * STAGE_BYTES;
	uint64_t *bars = reinterpret_cast<uint64_t *>(smem + a.tile_bar_off) + (size_t) warp * NSTAGE;

	if (lane == 0) {
#pragma unroll
		for (int s = 0; s < NSTAGE; s++) mbar_init(smem_u32(&bars[s]), 1);
	}
	stage_blob(smem, a.blob, a.blob_bytes, &blob_bar);   /* fences + syncs the inits too */

	TableSmem<E> ts;
	ts.tbl = reinterpret_cast<const E *>(smem);
	ts.pitch = a.pitch / (uint32_t) sizeof(E);
	const uint8_t *is_end = smem + a.is_end_off;

	const uint32_t ntiles = (uint32_t) ((a.n + 31) >> 5);
	const uint32_t gw = blockIdx.x * nwarps + warp;
	const uint32_t GW = gridDim.x * nwarps;
	if (gw >= ntiles) return;
	const uint32_t nst = (uint32_t) ((a.len + CH - 1) / CH);      /* stages per tile */

	/* swizzle: physical 16B chunk = logical ^ f(row) (CU_TENSOR_MAP_SWIZZLE_{32,64,128}B) */
	const uint32_t swz = (CH == 128) ? (lane & 7u) : (CH == 64) ? ((lane >> 1) & 3u) : ((lane >> 2) & 1u);
	const uint32_t row_off = lane * CH;
	const uint32_t stage_a = smem_u32(stage_base);
	const uint32_t bars_a = smem_u32(bars);

	/* Two cursors over this warp's sequence of (tile, stage) pairs: `ic` is the next pair
	 * to fetch, `cc` the next to consume; ic runs NSTAGE pairs ahead.  All counters are
	 * warp-uniform and advanced incrementally (no divisions in the loop). */
	struct Cursor { uint32_t tile, sidx, slot, phase; };
	Cursor cc = { gw, 0u, 0u, 0u }, ic = cc;
	auto advance = [&](Cursor &c) {
		if (++c.sidx == nst) { c.sidx = 0; c.tile += GW; }
		if (++c.slot == (uint32_t) NSTAGE) { c.slot = 0; c.phase ^= 1u; }
	};
	auto issue = [&]() {
		if (ic.tile < ntiles) {
			if (lane == 0) {
				const uint32_t bar = bars_a + ic.slot * 8u;
				mbar_expect_tx(bar, STAGE_BYTES);
				tma_tile_g2s(stage_a + ic.slot * STAGE_BYTES, &tmap, ic.sidx * CH, ic.tile << 5, bar);
			}
			advance(ic);
		}
	};
#pragma unroll
	for (int s = 0; s < NSTAGE; s++) issue();

	uint32_t st = a.start;
	uint64_t pos = 0;
	bool died = false;
	while (cc.tile < ntiles) {
		const uint64_t row = ((uint64_t) cc.tile << 5) + lane;
		if (cc.sidx == 0) { st = a.start; pos = 0; died = false; }

		mbar_wait(bars_a + cc.slot * 8u, cc.phase);

		const uint8_t *srow = stage_base + cc.slot * STAGE_BYTES + row_off;
		const uint64_t remain = a.len - (uint64_t) cc.sidx * CH;  /* bytes of this input left */
		if (!died && row < a.n) {
			if (remain >= CH) {
#pragma unroll
				for (int v = 0; v < NVEC; v++) {
					const uint4 x = *reinterpret_cast<const uint4 *>(srow + (((uint32_t) v ^ swz) << 4));
					const uint32_t entry = st;
					STEP4(ts, st, x.x); STEP4(ts, st, x.y); STEP4(ts, st, x.z); STEP4(ts, st, x.w);
					if (HAS_DEAD && st == a.dead) {
						/* a byte of this 16-byte chunk had no edge: re-walk it to find which */
						st = entry;
						for (uint32_t k = 0; k < 16; k++) {
							const uint32_t b = srow[((((uint32_t) v ^ swz) << 4) | k)];
							const uint32_t nx = ts.step(st, b);
							if (nx == a.dead) { died = true; pos += (uint64_t) k; break; }
							st = nx;
						}
						break;
					}
					pos += 16;
				}
			} else {
				for (uint32_t k = 0; k < (uint32_t) remain; k++) {
					const uint32_t b = srow[((((k >> 4) ^ swz) << 4) | (k & 15u))];
					const uint32_t nx = ts.step(st, b);
					if (HAS_DEAD && nx == a.dead) { died = true; break; }
					st = nx;
					pos++;
				}
			}
		}
		if (cc.sidx == nst - 1 && row < a.n) {
			const int32_t ret = (!died && is_end[st]) ? 1 : 0;
			store_result(a, row, ret, st, pos);
		}
		__syncwarp();       /* every lane has finished reading this slot */
		advance(cc);
		issue();            /* refill the slot just freed */
	}
}


/* ------------------------------------------------------------------ K-RANGE x TILE ------ */

/*
 * The k-range walk (ALU byte classification, one table read per 4 bytes) fed by the TMA tile ring
 * of k1_tile_kernel instead of per-lane 256-bit loads.  Why: with one lane per input every lane
 * streams its own 1 KiB row, and that access pattern -- whatever the load width or depth -- tops out
 * at 5.8 TB/s on B200 (tools/micro/membench.cu: 32 / 64 / 128 B per lane per step, 1-4 steps in
 * flight: 5.83 TB/s; 4 lanes reading 128 contiguous bytes of a row: 6.6 TB/s; 8 lanes x 256 B: 7.1;
 * fully coalesced: 7.2).  Here a warp's tile is 32 rows x CH bytes fetched by ONE 2-D TMA request
 * (CH contiguous bytes per row, hardware swizzle so that the lanes' 16-byte shared-memory reads do
 * not conflict), completed on an mbarrier; the lanes then walk their own row out of shared memory.
 * Fixed-stride, 16-byte aligned batches only (as k1_tile_kernel).
 * Shared memory: [k-range blob, padded to 1024][per warp: NSTAGE stages of 32 x CH bytes][mbarriers]
 */
template <bool HAS_DEAD, int RNG, int CH, int NSTAGE>
__global__ void __launch_bounds__(1024, 1)
k1_krange_tile_kernel(const K1Args a, const __grid_constant__ CUtensorMap tmap)
{
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint64_t blob_bar;
	__shared__ uint32_t main_done;
	constexpr uint32_t STAGE_BYTES = 32u * CH;
	constexpr int NVEC = CH / 16;

	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	const uint32_t nwarps = blockDim.x >> 5;
	uint8_t *stage_base = smem + a.tile_stage_off + (size_t) warp * NSTAGE * STAGE_BYTES;
	uint64_t *bars = reinterpret_cast<uint64_t *>(smem + a.tile_bar_off) + (size_t) warp * NSTAGE;

	if (lane == 0) {
#pragma unroll
		for (int s = 0; s < NSTAGE; s++) mbar_init(smem_u32(&bars[s]), 1);
	}
	if (threadIdx.x == 0) main_done = 0u;
	stage_blob(smem, a.kblob, a.kblob_bytes, &blob_bar);   /* fences + syncs the inits too */

	const uint8_t *L0 = smem;
	const uint8_t *t1 = smem + a.k1_off;
	const uint8_t *is_end = smem + a.kend_off;
	const uint32_t p1 = a.k1pitch;
	const uint32_t kt_half = (smem_u32(smem) + 256u) >> 1;
#define STEP1(st, b) ((uint32_t) t1[(st) * p1 + L0[(b)]])
#define KR_WORD(w)                                                                                   \
	do {                                                                                             \
		const uint32_t l_ = (w) & 0x7F7F7F7Fu;                                                       \
		const uint32_t h0_ = ((w) ^ a.kr_hxor[0]) & 0x80808080u;                                     \
		const uint32_t h1_ = RNG == 1 ? h0_ : (((w) ^ a.kr_hxor[1]) & 0x80808080u);                  \
		const uint32_t i0_ = (l_ + a.kr_add_lo[0]) & ~(l_ + a.kr_add_hi[0]) & h0_;                   \
		const uint32_t i1_ = (l_ + a.kr_add_lo[1]) & ~(l_ + a.kr_add_hi[1]) & h1_;                   \
		const uint32_t acc_ = __dp4a(i1_, 0x80200802u, __dp4a(i0_, 0x40100401u, kt_half));           \
		asm("ld.shared.u8 %0, [%1];" : "=r"(st) : "r"(st + acc_ + acc_));                            \
	} while (0)

	const uint32_t ntiles = (uint32_t) ((a.n + 31) >> 5);
	const uint32_t nst = (uint32_t) ((a.len + CH - 1) / CH);      /* stages per tile */

	/* Tile schedule.  A lane's walk is one dependent chain, so a warp's time per tile is fixed and an SM
	 * is full at `nmain` warps; static rounds of grid x nmain tiles then leave a ragged last round (config
	 * 2: 18.45 rounds cost 19).  When the host finds it pays, the CTA carries extra warps that sleep
	 * through the first `Rm` rounds and the tiles of the last two rounds are dealt to ALL warps at once,
	 * so every SM stays saturated to the end:
	 *   round r <  Rm : main warp g takes tile g + r x GWm
	 *   round r >= Rm : warp with all-warp number ga takes tile Rm x GWm + ga + (r - Rm) x GWa
	 * (main warps are numbered first among all warps, so leftovers go to warps that are awake anyway). */
	const uint32_t nmain = a.tile_main_warps ? a.tile_main_warps : nwarps;
	const uint32_t Rm = a.tile_main_warps ? a.tile_main_rounds : 0xFFFFFFFFu;
	const bool is_main = warp < nmain;
	const uint32_t GWm = gridDim.x * nmain, GWa = gridDim.x * nwarps;
	const uint32_t gwm = blockIdx.x * nmain + warp;
	const uint32_t gwa = is_main ? gwm : GWm + blockIdx.x * (nwarps - nmain) + (warp - nmain);
	const uint32_t F0 = a.tile_main_warps ? Rm * GWm : 0u;
	auto tile_of = [&](uint32_t r) -> uint32_t {
		return r < Rm ? gwm + r * GWm : F0 + gwa + (r - Rm) * GWa;
	};

	/* swizzle: physical 16B chunk = logical ^ f(row) (CU_TENSOR_MAP_SWIZZLE_{32,64,128}B) */
	const uint32_t swz = (CH == 128) ? (lane & 7u) : (CH == 64) ? ((lane >> 1) & 3u) : ((lane >> 2) & 1u);
	const uint32_t row_off = lane * CH;
	const uint32_t stage_a = smem_u32(stage_base);
	const uint32_t bars_a = smem_u32(bars);

	struct Cursor { uint32_t round, tile, sidx, slot, phase; };
	const uint32_t r0 = (is_main || Rm == 0xFFFFFFFFu) ? 0u : Rm;
	Cursor cc = { r0, tile_of(r0), 0u, 0u, 0u }, ic = cc;   /* tile >= ntiles: nothing to do, but stay for the signal */
	auto advance = [&](Cursor &c) {
		if (++c.sidx == nst) { c.sidx = 0; c.round++; c.tile = tile_of(c.round); }
		if (++c.slot == (uint32_t) NSTAGE) { c.slot = 0; c.phase ^= 1u; }
	};
	auto issue = [&]() {
		if (ic.tile < ntiles) {
			if (lane == 0) {
				const uint32_t bar = bars_a + ic.slot * 8u;
				mbar_expect_tx(bar, STAGE_BYTES);
				tma_tile_g2s(stage_a + ic.slot * STAGE_BYTES, &tmap, ic.sidx * CH, ic.tile << 5, bar);
			}
			advance(ic);
		}
	};
	if (!is_main && r0 != 0u) {
		/* sleep until half of this CTA's main warps have left their main rounds */
		while (*reinterpret_cast<volatile uint32_t *>(&main_done) * 2u < nmain) __nanosleep(2000);
	}
	bool announced = !is_main || Rm == 0xFFFFFFFFu || Rm == 0u;
#pragma unroll
	for (int s = 0; s < NSTAGE; s++) issue();

	uint32_t st = a.start;
	uint64_t pos = 0;
	bool died = false;
	while (cc.tile < ntiles) {
		const uint64_t row = ((uint64_t) cc.tile << 5) + lane;
		if (cc.sidx == 0) { st = a.start; pos = 0; died = false; }

		mbar_wait(bars_a + cc.slot * 8u, cc.phase);

		const uint8_t *srow = stage_base + cc.slot * STAGE_BYTES + row_off;
		const uint64_t remain = a.len - (uint64_t) cc.sidx * CH;  /* bytes of this input left */
		if (!died && row < a.n) {
			if (remain >= CH) {
#pragma unroll
				for (int v = 0; v < NVEC; v++) {
					const uint4 x = *reinterpret_cast<const uint4 *>(srow + (((uint32_t) v ^ swz) << 4));
					const uint32_t entry = st;
					KR_WORD(x.x); KR_WORD(x.y); KR_WORD(x.z); KR_WORD(x.w);
					if (HAS_DEAD && st == a.dead) {
						/* a byte of this 16-byte chunk had no edge: re-walk it to find which */
						st = entry;
						for (uint32_t k = 0; k < 16; k++) {
							const uint32_t b = srow[((((uint32_t) v ^ swz) << 4) | k)];
							const uint32_t nx = STEP1(st, b);
							if (nx == a.dead) { died = true; pos += (uint64_t) k; break; }
							st = nx;
						}
						break;
					}
					pos += 16;
				}
			} else {
				for (uint32_t k = 0; k < (uint32_t) remain; k++) {
					const uint32_t b = srow[((((k >> 4) ^ swz) << 4) | (k & 15u))];
					const uint32_t nx = STEP1(st, b);
					if (HAS_DEAD && nx == a.dead) { died = true; break; }
					st = nx;
					pos++;
				}
			}
		}
		if (cc.sidx == nst - 1 && row < a.n) {
			const int32_t ret = (!died && is_end[st]) ? 1 : 0;
			store_result(a, row, ret, st, pos);
		}
		__syncwarp();       /* every lane has finished reading this slot */
		advance(cc);
		if (!announced && cc.round >= Rm) {
			announced = true;
			if (lane == 0) atomicAdd(&main_done, 1u);
		}
		issue();            /* refill the slot just freed */
	}
	if (!announced && lane == 0) atomicAdd(&main_done, 1u);   /* a main warp that had no tile at all */
	signal_done(a);
#undef KR_WORD
#undef STEP1
}

/* ------------------------------------------------------------------ host side -------- */

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_fn
get_encode_tiled()
{
	static encode_tiled_fn fn = nullptr;
	static std::once_flag once;
	std::call_once(once, [] {
		void *p = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
		    qres == cudaDriverEntryPointSuccess) {
			fn = reinterpret_cast<encode_tiled_fn>(p);
		}
	});
	return fn;
}

int g_sm_count[64];
int g_smem_optin[64];

bool
device_props(int device, int *sms, int *smem)
{
	if (device < 0 || device >= 64) return false;
	if (g_sm_count[device] == 0) {
		int v = 0, m = 0;
		if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return false;
		if (cudaDeviceGetAttribute(&m, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return false;
		g_sm_count[device] = v;
		g_smem_optin[device] = m;
	}
	*sms = g_sm_count[device];
	*smem = g_smem_optin[device];
	return true;
}

template <typename K>
bool
set_smem(K kernel, size_t bytes)
{
	return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes) == cudaSuccess;
}

template <typename E, bool SMEM, bool HAS_DEAD, bool CLS>
int
launch_lane(const K1Args &a, int sms, size_t smem_bytes, int block, cudaStream_t stream)
{
	auto kern = k1_lane_kernel<E, SMEM, HAS_DEAD, CLS>;
	if (SMEM && !set_smem(kern, smem_bytes)) {
		set_error("k1_lane: cannot opt in to %zu bytes of shared memory", smem_bytes);
		errno = EIO;
		return -1;
	}
	int per_sm = 1;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, SMEM ? smem_bytes : 0) != cudaSuccess || per_sm < 1) {
		per_sm = 1;
	}
	uint64_t want = (a.n + (uint64_t) block - 1) / (uint64_t) block;
	uint64_t grid = (uint64_t) sms * (uint64_t) per_sm;
	if (want < grid) grid = want;
	if (grid == 0) grid = 1;
	kern<<<(unsigned) grid, block, SMEM ? smem_bytes : 0, stream>>>(a);
	count_launch();
	FSMB_CUDA(cudaGetLastError(), return -1);
	return 0;
}

template <int K, bool HAS_DEAD, int RNG>
int
launch_kstride(const K1Args &a, int sms, cudaStream_t stream)
{
	void (*kern)(const K1Args);
	if (RNG == 1) kern = k1_krange_kernel<HAS_DEAD, 1>;
	else if (RNG == 2) kern = k1_krange_kernel<HAS_DEAD, 2>;
	else kern = k1_kstride_kernel<K, HAS_DEAD>;
	const size_t smem_bytes = (a.kblob_bytes + 127u) & ~127u;
	if (!set_smem(kern, smem_bytes)) {
		set_error("k1_kstride: cannot opt in to %zu bytes of shared memory", smem_bytes);
		errno = EIO;
		return -1;
	}
	int block = 1024;
	if (const char *e = getenv("FSM_B200_KSTRIDE_BLOCK")) {      /* tuning knob */
		const int v = atoi(e);
		if (v >= 64 && v <= 1024 && (v % 32) == 0) block = v;
	}
	int per_sm = 1;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, smem_bytes) != cudaSuccess || per_sm < 1) per_sm = 1;
	uint64_t want = (a.n + (uint64_t) block - 1) / (uint64_t) block;
	uint64_t grid = (uint64_t) sms * (uint64_t) per_sm;
	if (want < grid) grid = want;
	if (grid == 0) grid = 1;
	kern<<<(unsigned) grid, block, smem_bytes, stream>>>(a);
	count_launch();
	FSMB_CUDA(cudaGetLastError(), return -1);
	return 0;
}

template <typename E, bool HAS_DEAD, int CH, int NSTAGE>
int
launch_tile(K1Args a, int sms, int smem_optin, cudaStream_t stream)
{
	auto kern = k1_tile_kernel<E, HAS_DEAD, CH, NSTAGE>;
	encode_tiled_fn enc = get_encode_tiled();
	if (enc == nullptr) {
		set_error("k1_tile: cuTensorMapEncodeTiled unavailable");
		errno = EIO;
		return -1;
	}
	const uint32_t blob_pad = (a.blob_bytes + 1023u) & ~1023u;
	const uint32_t per_warp = (uint32_t) NSTAGE * 32u * CH;
	int nwarps = (int) (((uint32_t) smem_optin - blob_pad - 1024u) / (per_warp + 8u * NSTAGE));
	if (nwarps > 32) nwarps = 32;
	if (nwarps < 1) {
		set_error("k1_tile: table too large for the tiled variant");
		errno = ENOTSUP;
		return -1;
	}
	const uint64_t ntiles = (a.n + 31) >> 5;
	a.tile_stage_off = blob_pad;
	a.tile_bar_off = blob_pad + (uint32_t) nwarps * per_warp;
	const size_t smem_bytes = (size_t) a.tile_bar_off + (size_t) nwarps * NSTAGE * 8u;
	if (!set_smem(kern, smem_bytes)) {
		set_error("k1_tile: cannot opt in to %zu bytes of shared memory", smem_bytes);
		errno = EIO;
		return -1;
	}

	CUtensorMap tmap;
	const cuuint64_t gdim[2] = { (cuuint64_t) a.stride, (cuuint64_t) a.n };
	const cuuint64_t gstr[1] = { (cuuint64_t) a.stride };
	const cuuint32_t box[2] = { (cuuint32_t) CH, 32u };
	const cuuint32_t estr[2] = { 1u, 1u };
	const CUtensorMapSwizzle sw = (CH == 128) ? CU_TENSOR_MAP_SWIZZLE_128B
	    : (CH == 64) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
	CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(a.base), gdim, gstr,
	    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
	    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		set_error("k1_tile: cuTensorMapEncodeTiled failed (%d)", (int) r);
		errno = EIO;
		return -1;
	}
	uint64_t grid = (ntiles + (uint64_t) nwarps - 1) / (uint64_t) nwarps;
	if (grid > (uint64_t) sms) grid = (uint64_t) sms;
	if (grid == 0) grid = 1;
	kern<<<(unsigned) grid, nwarps * 32, smem_bytes, stream>>>(a, tmap);
	count_launch();
	FSMB_CUDA(cudaGetLastError(), return -1);
	return 0;
}

template <bool HAS_DEAD, int RNG, int CH, int NSTAGE>
int
launch_krange_tile(K1Args a, int sms, int smem_optin, cudaStream_t stream)
{
	auto kern = k1_krange_tile_kernel<HAS_DEAD, RNG, CH, NSTAGE>;
	encode_tiled_fn enc = get_encode_tiled();
	if (enc == nullptr) {
		set_error("k1_krange_tile: cuTensorMapEncodeTiled unavailable");
		errno = EIO;
		return -1;
	}
	const uint32_t blob_pad = (a.kblob_bytes + 1023u) & ~1023u;
	const uint32_t per_warp = (uint32_t) NSTAGE * 32u * CH;
	int nwarps = (int) (((uint32_t) smem_optin - blob_pad - 1024u) / (per_warp + 8u * NSTAGE));
	if (nwarps > 32) nwarps = 32;
	/* measured on B200 (config 2, CH = 128, 2 stages): throughput grows linearly up to 12 warps per SM
	 * (each warp is bound by its own lookup chain) and is flat to slightly worse beyond (profiles/) */
	if (nwarps > 12) nwarps = 12;
	if (const char *e = getenv("FSM_B200_KRTILE_WARPS")) {         /* tuning knob */
		const int v = atoi(e);
		if (v >= 1 && v <= 32 && (size_t) blob_pad + 1024u + (size_t) v * (per_warp + 8u * NSTAGE) <= (size_t) smem_optin) nwarps = v;
	}
	if (nwarps < 1) {
		set_error("k1_krange_tile: table too large");
		errno = ENOTSUP;
		return -1;
	}
	const uint64_t ntiles = (a.n + 31) >> 5;
	uint64_t grid = (ntiles + (uint64_t) nwarps - 1) / (uint64_t) nwarps;
	if (grid > (uint64_t) sms) grid = (uint64_t) sms;
	if (grid == 0) grid = 1;
	/* ragged last round: carry sleeping extra warps and deal the last two rounds' tiles to all warps at
	 * once when that is predicted to be quicker (see the kernel's tile schedule).  Cost model, in units
	 * of one warp's tile time: an SM runs `nwarps` walks at full speed and is issue-bound beyond that;
	 * oversubscription costs ~5% more (measured: 16 warps per SM run at 0.95 of 12). */
	a.tile_main_warps = 0; a.tile_main_rounds = 0;
	{
		const char *e = getenv("FSM_B200_KRTILE_TAIL");
		const bool allow = e == nullptr || atoi(e) != 0;
		const uint64_t GWm = grid * (uint64_t) nwarps;
		const uint64_t R = (ntiles + GWm - 1) / GWm;
		const int max_warps = (int) std::min<uint64_t>(32, ((uint64_t) smem_optin - blob_pad - 1024u) / (per_warp + 8u * NSTAGE));
		if (allow && grid == (uint64_t) sms && R >= 2 && R < (1u << 20) && max_warps > nwarps) {
			const uint64_t Fn = ntiles - (R - 2) * GWm;
			const int need = (int) ((Fn + grid - 1) / grid);             /* warps per CTA for one final round */
			if (need <= max_warps) {
				const double cost = (double) (R - 2) + 1.05 * std::max(1.0, (double) Fn / (double) grid / (double) nwarps);
				if (cost < (double) R - 0.05) {
					a.tile_main_warps = (uint32_t) nwarps;
					a.tile_main_rounds = (uint32_t) (R - 2);
					nwarps = std::max(need, nwarps);
				}
			}
		}
	}
	a.tile_stage_off = blob_pad;
	a.tile_bar_off = blob_pad + (uint32_t) nwarps * per_warp;
	const size_t smem_bytes = (size_t) a.tile_bar_off + (size_t) nwarps * NSTAGE * 8u;
	if (!set_smem(kern, smem_bytes)) {
		set_error("k1_krange_tile: cannot opt in to %zu bytes of shared memory", smem_bytes);
		errno = EIO;
		return -1;
	}
	CUtensorMap tmap;
	const cuuint64_t gdim[2] = { (cuuint64_t) a.stride, (cuuint64_t) a.n };
	const cuuint64_t gstr[1] = { (cuuint64_t) a.stride };
	const cuuint32_t box[2] = { (cuuint32_t) CH, 32u };
	const cuuint32_t estr[2] = { 1u, 1u };
	const CUtensorMapSwizzle sw = (CH == 128) ? CU_TENSOR_MAP_SWIZZLE_128B
	    : (CH == 64) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
	CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(a.base), gdim, gstr,
	    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
	    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		set_error("k1_krange_tile: cuTensorMapEncodeTiled failed (%d)", (int) r);
		errno = EIO;
		return -1;
	}
	kern<<<(unsigned) grid, nwarps * 32, smem_bytes, stream>>>(a, tmap);
	count_launch();
	FSMB_CUDA(cudaGetLastError(), return -1);
	return 0;
}

int g_variant = 0;

struct K1SignalArgs { uint32_t *flags[8]; uint32_t n, value; };

/* the completion signal of a rank that had nothing to scan (see signal_done) */
__global__ void
k1_signal_only_kernel(const K1SignalArgs a)
{
	if (threadIdx.x == 0) {
		for (uint32_t r = 0; r < a.n; r++) *reinterpret_cast<volatile uint32_t *>(a.flags[r]) = a.value;
		__threadfence_system();
	}
}

int
dispatch_kstride(const fsm_b200_dfa *dfa, const K1Args &a0, int sms, cudaStream_t stream, bool tile_ok = false)
{
	const bool dead = !dfa->complete;
	uint32_t rng = (dfa->kstride == 4 && dfa->d_rblob != nullptr) ? dfa->krange : 0u;
	if (getenv("FSM_B200_KSTRIDE_LUT") != nullptr) rng = 0;        /* tuning knob: class LUTs even when ranges exist */
	if (rng != 0) {
		K1Args a = a0;
		a.kblob = static_cast<const uint8_t *>(dfa->d_rblob);
		a.kblob_bytes = dfa->rblob_bytes; a.k1_off = dfa->r_k1_off; a.kend_off = dfa->r_kend_off;
		a.kr_prefetch = 0;
		if (const char *e = getenv("FSM_B200_KRANGE_PREFETCH")) {      /* tuning knob: L2 prefetch distance, 0 = off */
			const int v = atoi(e);
			if (v >= 0 && v <= 65536 && (v % 128) == 0) a.kr_prefetch = (uint32_t) v;
		}
		/* fixed-stride, 16-byte aligned batches of inputs at least a tile wide: TMA tile ring (the DRAM
		 * access pattern of one lane per row caps at 5.8 TB/s, see k1_krange_tile_kernel) */
		int tile = 128;
		if (const char *e = getenv("FSM_B200_KRANGE_TILE")) tile = atoi(e);          /* tuning knob: 0 = per-lane loads, 64, 128 */
		if (tile_ok && a.len >= 256 && a.entry == nullptr && a.n_dev == nullptr && (tile == 128 || tile == 64)) {
			int smem_optin = 0, sms2 = 0;
			if (device_props(dfa->device, &sms2, &smem_optin)) {
				int stages = 2;
				if (const char *e = getenv("FSM_B200_KRTILE_STAGES")) { const int v = atoi(e); if (v == 2 || v == 3 || v == 4) stages = v; }
#define KRT(RNGV, CHV, NSV) (dead ? launch_krange_tile<true, RNGV, CHV, NSV>(a, sms, smem_optin, stream) : launch_krange_tile<false, RNGV, CHV, NSV>(a, sms, smem_optin, stream))
				if (tile == 128) {
					if (rng == 1) return stages == 2 ? KRT(1, 128, 2) : stages == 3 ? KRT(1, 128, 3) : KRT(1, 128, 4);
					return stages == 2 ? KRT(2, 128, 2) : stages == 3 ? KRT(2, 128, 3) : KRT(2, 128, 4);
				}
				if (rng == 1) return stages == 2 ? KRT(1, 64, 2) : stages == 3 ? KRT(1, 64, 3) : KRT(1, 64, 4);
				return stages == 2 ? KRT(2, 64, 2) : stages == 3 ? KRT(2, 64, 3) : KRT(2, 64, 4);
#undef KRT
			}
		}
		if (rng == 1) return dead ? launch_kstride<4, true, 1>(a, sms, stream) : launch_kstride<4, false, 1>(a, sms, stream);
		return dead ? launch_kstride<4, true, 2>(a, sms, stream) : launch_kstride<4, false, 2>(a, sms, stream);
	}
	if (dfa->kstride == 4) return dead ? launch_kstride<4, true, 0>(a0, sms, stream) : launch_kstride<4, false, 0>(a0, sms, stream);
	return dead ? launch_kstride<2, true, 0>(a0, sms, stream) : launch_kstride<2, false, 0>(a0, sms, stream);
}

} // namespace

namespace fsmb200 {

bool
k1_tile_eligible(const fsm_b200_dfa *dfa, const uint8_t *d_base, const uint64_t *d_offsets,
	uint64_t stride, uint64_t len, size_t n)
{
	return dfa->smem_resident && dfa->nclasses == 0 && d_offsets == nullptr && n > 0 && len > 0 &&
	    (reinterpret_cast<uintptr_t>(d_base) & 15u) == 0 && (stride & 15u) == 0 &&
	    stride >= len && stride < (1ull << 32) && n < (1ull << 31) && dfa->entry_bytes <= 2;
}

/* what the TMA tile ring needs of the BATCH (any table): fixed stride, 16-byte aligned */
static bool
k1_tile_base_ok(const uint8_t *d_base, const uint64_t *d_offsets, uint64_t stride, uint64_t len, size_t n)
{
	return d_offsets == nullptr && n > 0 && len > 0 && (reinterpret_cast<uintptr_t>(d_base) & 15u) == 0 && (stride & 15u) == 0 &&
	    stride >= len && stride < (1ull << 32) && n < (1ull << 31);
}

static void
fill_args(K1Args &a, const fsm_b200_dfa *dfa)
{
	memset(&a, 0, sizeof a);
	a.blob = static_cast<const uint8_t *>(dfa->d_blob);
	a.blob_bytes = (uint32_t) dfa->blob_bytes;
	a.is_end_off = dfa->is_end_off;
	a.cls_off = dfa->cls_off;
	a.pitch = dfa->pitch; a.start = dfa->start; a.dead = dfa->dead;
	a.absorb = (dfa->has_absorbing && getenv("FSM_B200_NO_ABSORB_SKIP") == nullptr) ? dfa->d_absorb : nullptr;
	a.kblob = static_cast<const uint8_t *>(dfa->d_kblob);
	a.kblob_bytes = dfa->kblob_bytes; a.kpitch = dfa->kpitch; a.k1pitch = dfa->k1pitch;
	a.k1_off = dfa->k1_off; a.kend_off = dfa->kend_off; a.klut_off = dfa->klut_off;
	for (int k = 0; k < 2; k++) { a.kr_add_lo[k] = dfa->kr_add_lo[k]; a.kr_add_hi[k] = dfa->kr_add_hi[k]; a.kr_hxor[k] = dfa->kr_hxor[k]; }
}

template <typename E, bool SMEM, bool HAS_DEAD, bool CLS>
int
launch_ragged(const K1Args &a, int sms, size_t smem_bytes, int block, cudaStream_t stream)
{
	auto kern = k1_ragged_kernel<E, SMEM, HAS_DEAD, CLS>;
	if (block > 768) block = 768;
	if (SMEM && !set_smem(kern, smem_bytes)) {
		set_error("k1_ragged: cannot opt in to %zu bytes of shared memory", smem_bytes);
		errno = EIO;
		return -1;
	}
	int per_sm = 1;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, SMEM ? smem_bytes : 0) != cudaSuccess || per_sm < 1) {
		per_sm = 1;
	}
	uint64_t want = (a.n + (uint64_t) block - 1) / (uint64_t) block;
	uint64_t grid = (uint64_t) sms * (uint64_t) per_sm;
	if (want < grid) grid = want;
	if (grid == 0) grid = 1;
	kern<<<(unsigned) grid, block, SMEM ? smem_bytes : 0, stream>>>(a);
	count_launch();
	FSMB_CUDA(cudaGetLastError(), return -1);
	return 0;
}

template <typename E, bool SMEM>
static int
dispatch_lane2(const fsm_b200_dfa *dfa, const K1Args &a, int sms, size_t smem_bytes, int block, cudaStream_t stream)
{
	const bool dead = !dfa->complete, cls = dfa->nclasses != 0;
	if (a.offsets != nullptr && !a.prefer_lane && getenv("FSM_B200_NO_RAGGED") == nullptr) {
		if (dead) {
			return cls ? launch_ragged<E, SMEM, true, true>(a, sms, smem_bytes, block, stream)
			           : launch_ragged<E, SMEM, true, false>(a, sms, smem_bytes, block, stream);
		}
		return cls ? launch_ragged<E, SMEM, false, true>(a, sms, smem_bytes, block, stream)
		           : launch_ragged<E, SMEM, false, false>(a, sms, smem_bytes, block, stream);
	}
	if (dead) {
		return cls ? launch_lane<E, SMEM, true, true>(a, sms, smem_bytes, block, stream)
		           : launch_lane<E, SMEM, true, false>(a, sms, smem_bytes, block, stream);
	}
	return cls ? launch_lane<E, SMEM, false, true>(a, sms, smem_bytes, block, stream)
	           : launch_lane<E, SMEM, false, false>(a, sms, smem_bytes, block, stream);
}

static int
dispatch_lane(const fsm_b200_dfa *dfa, const K1Args &a, int sms, cudaStream_t stream)
{
	int block = 1024;
	const char *e = getenv("FSM_B200_LANE_BLOCK");
	if (e != nullptr) {
		int v = atoi(e);
		if (v >= 32 && v <= 1024 && (v % 32) == 0) block = v;
	}
	if (dfa->smem_resident) {
		const size_t smem_bytes = (a.blob_bytes + 127u) & ~127u;
		if (dfa->entry_bytes == 1) return dispatch_lane2<uint8_t, true>(dfa, a, sms, smem_bytes, block, stream);
		if (dfa->entry_bytes == 2) return dispatch_lane2<uint16_t, true>(dfa, a, sms, smem_bytes, block, stream);
		return dispatch_lane2<uint32_t, true>(dfa, a, sms, smem_bytes, block, stream);
	}
	if (block > 256 && e == nullptr) block = 256;
	if (dfa->entry_bytes == 1) return dispatch_lane2<uint8_t, false>(dfa, a, sms, 0, block, stream);
	if (dfa->entry_bytes == 2) return dispatch_lane2<uint16_t, false>(dfa, a, sms, 0, block, stream);
	return dispatch_lane2<uint32_t, false>(dfa, a, sms, 0, block, stream);
}

int
k1_launch_jobs(const fsm_b200_dfa *dfa, const uint8_t *d_base, const uint64_t *d_begs,
	const uint64_t *d_ends, const uint32_t *d_entry, size_t n, const uint32_t *d_n,
	fsm_b200_result *d_out, cudaStream_t stream)
{
	if (n == 0) return 0;
	int sms = 0, smem_optin = 0;
	if (!device_props(dfa->device, &sms, &smem_optin)) {
		set_error("k1: cannot query device %d", dfa->device);
		errno = EIO;
		return -1;
	}
	K1Args a;
	fill_args(a, dfa);
	a.base = d_base; a.offsets = d_begs; a.ends = d_ends; a.entry = d_entry; a.n = n; a.n_dev = d_n; a.out = d_out;
	a.prefer_lane = 1;
	if (dfa->kstride != 0 && getenv("FSM_B200_STREAM_NO_KSTRIDE") == nullptr) {
		return dispatch_kstride(dfa, a, sms, stream);
	}
	return dispatch_lane(dfa, a, sms, stream);
}

int
k1_launch(const fsm_b200_dfa *dfa, const uint8_t *d_base, const uint64_t *d_offsets,
	uint64_t stride, uint64_t len, size_t n, fsm_b200_result *d_out, cudaStream_t stream, int variant,
	fsm_b200_result *const *peer_outs, int npeers, int peer_compact,
	uint32_t *sig_counter, uint32_t *const *sig_flags, uint32_t sig_value)
{
	if (n == 0) {
		/* an empty shard still has to publish its completion flag, or a peer polling it waits forever */
		if (sig_counter != nullptr && sig_flags != nullptr && npeers >= 0 && npeers <= 7) {
			K1SignalArgs sa;
			for (int r = 0; r <= npeers; r++) sa.flags[r] = sig_flags[r];
			sa.n = (uint32_t) npeers + 1u; sa.value = sig_value;
			k1_signal_only_kernel<<<1, 32, 0, stream>>>(sa);
			count_launch();
			FSMB_CUDA(cudaGetLastError(), return -1);
		}
		return 0;
	}
	int sms = 0, smem_optin = 0;
	if (!device_props(dfa->device, &sms, &smem_optin)) {
		set_error("k1: cannot query device %d", dfa->device);
		errno = EIO;
		return -1;
	}
	K1Args a;
	fill_args(a, dfa);
	a.base = d_base; a.offsets = d_offsets; a.stride = stride; a.len = len; a.n = n; a.out = d_out;
	if (npeers < 0 || npeers > 7) {
		set_error("k1: at most 7 peers (8 GPUs of one node)");
		errno = EINVAL;
		return -1;
	}
	for (int r = 0; r < npeers; r++) a.peer_out[r] = peer_outs[r];
	a.npeers = (uint32_t) npeers;
	a.peer_compact = peer_compact ? 1u : 0u;
	if (sig_counter != nullptr && sig_flags != nullptr) {
		a.sig_counter = sig_counter;
		for (int r = 0; r <= npeers; r++) a.sig_flags[r] = sig_flags[r];
		a.sig_value = sig_value;
	}

	if (variant == K1_AUTO) variant = g_variant;
	/* ragged batches on shared-memory-resident tables: the lines kernel (k1_lines.cu) */
	if (variant == K1_AUTO && d_offsets != nullptr && npeers == 0 && sig_counter == nullptr &&
	    k1_lines_eligible(dfa) && getenv("FSM_B200_K1_VARIANT") == nullptr) {
		return k1_lines_launch(dfa, d_base, d_offsets, 0, 0, n, d_out, nullptr, stream);
	}
	const bool tile_ok = k1_tile_eligible(dfa, d_base, d_offsets, stride, len, n);
	if (variant == K1_AUTO) {
		/* Measured on B200 (profiles/r1_k1_variants.jsonl): LANE >= TILE64 on both config-2
		 * distributions -- both saturate the L1TEX data pipe (ncu: l1tex 97 %), and LANE also
		 * takes ragged/unaligned batches.  FSM_B200_K1_VARIANT overrides. */
		variant = K1_LANE;
		/* fixed-stride batches on DFAs with few byte classes: one lookup per K bytes */
		if (dfa->kstride != 0 && d_offsets == nullptr) variant = K1_KSTRIDE;
		if (const char *e = getenv("FSM_B200_K1_VARIANT")) {
			const int v = atoi(e);
			if (v > K1_AUTO && v < K1_VARIANT_COUNT &&
			    (v == K1_LANE || (v == K1_KSTRIDE && dfa->kstride != 0) || (v != K1_KSTRIDE && tile_ok))) variant = v;
		}
	}
	const bool dead = !dfa->complete;

	if (variant == K1_LANE) {
		return dispatch_lane(dfa, a, sms, stream);
	}
	if (variant == K1_KSTRIDE) {
		if (dfa->kstride == 0) {
			set_error("k1: this DFA has no k-stride table (more than 16 byte classes or more than 256 rows)");
			errno = ENOTSUP;
			return -1;
		}
		return dispatch_kstride(dfa, a, sms, stream, k1_tile_base_ok(d_base, d_offsets, stride, len, n));
	}

	if (!tile_ok) {
		set_error("k1: tiled variant needs a shared-memory-resident table and a fixed-stride, 16-byte aligned batch");
		errno = ENOTSUP;
		return -1;
	}
#define TILE_CASE(V, CH, NS)                                                                       \
	if (variant == (V)) {                                                                          \
		if (dfa->entry_bytes == 1)                                                                 \
			return dead ? launch_tile<uint8_t, true, CH, NS>(a, sms, smem_optin, stream)           \
			            : launch_tile<uint8_t, false, CH, NS>(a, sms, smem_optin, stream);         \
		return dead ? launch_tile<uint16_t, true, CH, NS>(a, sms, smem_optin, stream)              \
		            : launch_tile<uint16_t, false, CH, NS>(a, sms, smem_optin, stream);            \
	}
	TILE_CASE(K1_TILE64, 64, 2)
	TILE_CASE(K1_TILE32, 32, 4)
	TILE_CASE(K1_TILE128, 128, 2)
	TILE_CASE(K1_TILE64x3, 64, 3)
#undef TILE_CASE
	set_error("k1: unknown variant %d", variant);
	errno = EINVAL;
	return -1;
}

} // namespace fsmb200

extern "C" int
fsm_b200_set_exec_variant(int variant)
{
	if (variant < 0 || variant >= K1_VARIANT_COUNT) {
		errno = EINVAL;
		return -1;
	}
	g_variant = variant;
	return 0;
}

extern "C" int
fsm_b200_get_exec_variant(void)
{
	return g_variant;
}
