/*
 * k1b_rep_tma.cuh -- the fused small-automaton form of K1b (k1b_rep.cuh) with the INPUT taken off the LSU
 * path: every warp fetches "sector i of my 32 chunks" as one 32-row x 32-byte 2-D TMA tile into a ring in
 * shared memory and the lanes read their own row with two LDS.128.
 *
 * Why: with three 256-bit loads in flight per lane the LDG form is L1TEX-bound (85 %), and 28 of those
 * points are the loads themselves -- a 256-bit load whose 32 lanes touch 32 different lines costs 21
 * data-pipe wavefronts per request against 32 for the 32 table lookups it feeds.  A TMA tile does not pass
 * the LSU data pipe at all; reading it back is 8 wavefronts per sector-warp (swizzled: conflict-free).
 *
 * What it costs: the ring needs 1 KiB per warp and stage, and three tiles in flight per warp (the depth the
 * LDG form needed) make 128 KiB per CTA -- which does not fit next to a replicated table with the bit-7
 * gap of k1b_rep.cuh (16 KiB per row).  So this form uses the COMPACT layout, 8 KiB per row:
 *
 *   address of entry (state, byte) of lane l = state << 13 | (byte >> 3) << 8 | (byte >> 2 & 1) << 7 | l << 2 | (byte & 3)
 *
 * Still bank = lane; the byte-dependent part is still ONE prmt.b32 per byte, {q.byte i, h.byte i, 0, 0} with
 * per word h = (w >> 3) & 0x1F1F1F1F and q = (w & 0x03030303) | ((w << 5) & 0x80808080) | lane bits: five
 * instructions per word instead of three (4.25 per byte instead of 3.75).
 *
 * MEASURED (round 2, profiles/r2_k1b_rep_tma_ncu_full.txt, r2_k1b_rep_knobs.jsonl): slower than the 256-bit-load
 * form -- 2 GiB of UTF-8 in 1.24 ms with 4 stages, 0.96 ms with 3, 0.59 ms with 2, against 0.50 ms.  The more
 * tiles are in flight the slower it gets; at 4 stages the TMA unit delivers one 32-byte box ROW per 5.3 cycles
 * per SM (14.1 K tiles per SM in 2.39 M cycles = 169 cycles per 32-row tile: 1.8 TB/s) and 65 % of all warp
 * samples sit on the full barrier -- skinny rows are the wrong shape for it (the 128-byte rows of
 * k1_krange_tile_kernel reach 6 TB/s), and wider rows would need 4 KiB per warp and stage.  Off by default (FSM_B200_REP_TMA=1 selects it); kept as the record of the
 * experiment, parity-tested like every other variant.
 *
 * Everything else -- chunking, prefix, distinct live images, chunk maps, warp / CTA folds, the final kernel --
 * is k1b_rep.cuh's.  Warps whose 32 chunks are not all full rows of the tensor (the last one) and unaligned
 * buffers read with 256-bit loads as before.
 */
#ifndef FSM_B200_K1B_REP_TMA_CUH
#define FSM_B200_K1B_REP_TMA_CUH

#include <cuda.h>

#include "k1b_rep.cuh"

namespace fsmb200 {

constexpr uint32_t REPC_ROW_SHIFT = 13;                  /* 8 KiB per table row */
constexpr uint32_t REPC_BARS_BYTES = 32u * 4u * 8u;      /* full barriers: [32 warps][4 stages] */
constexpr uint32_t REPC_HEAD_BYTES = REP_MAPS_BYTES + REPC_BARS_BYTES;   /* in front of the table */
constexpr uint32_t REPC_STAGE_BYTES = 32u * 32u;         /* one tile: 32 rows x 32 B */

__device__ __forceinline__ void
repc_tma_tile(uint32_t dst, const CUtensorMap *map, uint32_t x, uint32_t y, uint32_t bar)
{
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes"
	    " [%0], [%1, {%2, %3}], [%4];"
	    :: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}

__device__ __forceinline__ uint4
lds_u128(uint32_t addr)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
	return v;
}

#define REPC_STEP1(st, b) lds_u8(((st) << REPC_ROW_SHIFT) + (((uint32_t) (b) >> 3) << 8) + ((((uint32_t) (b) >> 2) & 1u) << 7) + ((uint32_t) (b) & 3u) + lane4)

#define REPC_STEP4(st, w)                                                      \
	do {                                                                       \
		const uint32_t h_ = ((w) >> 3) & 0x1F1F1F1Fu;                          \
		const uint32_t u_ = ((w) & 0x03030303u) | lanebits;                    \
		const uint32_t q_ = u_ | (((w) << 5) & 0x80808080u);                   \
		st = lds_u8(((st) << REPC_ROW_SHIFT) + prmt_sx(q_, h_, 0xCC40u));      \
		st = lds_u8(((st) << REPC_ROW_SHIFT) + prmt_sx(q_, h_, 0xCC51u));      \
		st = lds_u8(((st) << REPC_ROW_SHIFT) + prmt_sx(q_, h_, 0xCC62u));      \
		st = lds_u8(((st) << REPC_ROW_SHIFT) + prmt_sx(q_, h_, 0xCC73u));      \
	} while (0)

/* a.C, a.mis == 0, the tensor map: rows of C bytes, `nfull` of them (the chunks that lie wholly in the buffer) */
template <bool HAS_DEAD, int NST>
__global__ void __launch_bounds__(1024, 1)
k1b_rep_tma_kernel(const RepArgs a, const __grid_constant__ CUtensorMap tmap, const uint32_t nfull, const uint32_t ring_off)
{
	extern __shared__ __align__(1024) uint8_t dsm[];
	uint64_t *wm = reinterpret_cast<uint64_t *>(dsm);
	uint32_t *s_wdc = reinterpret_cast<uint32_t *>(dsm + REP_CMAPS_BYTES);
	uint8_t *s_wmap = dsm + REP_CMAPS_BYTES + 32u * 16u * 4u;
	uint8_t *s_wds = s_wmap + 32u * 16u;
	uint32_t *s_done = reinterpret_cast<uint32_t *>(s_wds + 32u * 16u);
	const uint32_t dsm_a = smem_u32(dsm);
	const uint32_t rep_base = (dsm_a + REPC_HEAD_BYTES + 8191u) & ~8191u;      /* shared address of the table */
	const uint32_t K = rep_base >> REPC_ROW_SHIFT;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint32_t bars_a = dsm_a + REP_MAPS_BYTES + warp * (4u * 8u);          /* this warp's full barriers */
	const uint32_t ring_a = dsm_a + ring_off + warp * (uint32_t) NST * REPC_STAGE_BYTES;

	if (threadIdx.x == 0) *s_done = 0;
	if (lane == 0) {
#pragma unroll
		for (int s = 0; s < NST; s++) mbar_init(bars_a + 8u * s, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	/* replicate the table (compact layout): word j of row s of lane l at s << 13 | (j >> 1) << 8 | (j & 1) << 7 | l << 2 */
	{
		const uint32_t *d32 = reinterpret_cast<const uint32_t *>(a.dense);
		const uint32_t total = a.ntable * 64u * 32u;
		const uint32_t k4 = K * 0x01010101u;
		for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
			const uint32_t l = i & 31u, j = (i >> 5) & 63u, s = i >> 11;
			const uint32_t v = __ldg(d32 + s * 64u + j) + k4;
			asm volatile("st.shared.u32 [%0], %1;" :: "r"(rep_base + ((s << REPC_ROW_SHIFT) | ((j >> 1) << 8) | ((j & 1u) << 7) | (l << 2))), "r"(v) : "memory");
		}
	}
	__syncthreads();

	const uint32_t gw = blockIdx.x * a.wpc + warp;
	if (warp >= a.wpc || gw >= a.nwarps) return;
	const uint32_t nactive = min(a.wpc, a.nwarps - blockIdx.x * a.wpc);
	const uint32_t lane4 = lane << 2;
	const uint32_t lanebits = lane4 * 0x01010101u;
	const uint32_t T = a.T;
	const uint32_t deadK = a.dead + K;
	const uint32_t c = gw * 32u + lane;
	const bool have = c < a.nchunks;
	const uint64_t beg = have ? rep_chunk_beg(a, c) : 0, end = have ? rep_chunk_end(a, c) : 0;
	/* all 32 chunks of this warp are full rows of the tensor: the body comes in by TMA tiles */
	const bool tiled = gw * 32u + 32u <= nfull;
	const uint32_t sw16 = ((lane >> 2) & 1u) << 4;             /* CU_TENSOR_MAP_SWIZZLE_32B: 16-byte chunk ^= row bit 2 */
	uint32_t uses = 0;                                         /* tiles this warp has consumed: slot = uses % NST, parity = uses / NST */

	uint64_t img = 0;
	uint64_t exit_of = 0xFEDCBA9876543210ull;
	uint32_t live = 0;
	uint32_t w = 0;
	if (have) {
		w = (uint32_t) min((uint64_t) REP_W, end - beg);
		const uint8_t *p = a.buf + beg;
		if (w == REP_W && (reinterpret_cast<uintptr_t>(p) & 31u) == 0) {
			uint32_t A[8], B[8];
			ld256(p, A);
			ld256(p + 32, B);
#pragma unroll 1
			for (uint32_t s = 0; s < T; s++) {
				uint32_t st = s + K;
#pragma unroll
				for (int k = 0; k < 8; k++) REPC_STEP4(st, A[k]);
#pragma unroll
				for (int k = 0; k < 8; k++) REPC_STEP4(st, B[k]);
				st = (HAS_DEAD && st == deadK) ? REP_DIED_PREFIX : st - K;
				img |= (uint64_t) st << (4u * s);
			}
		} else {
#pragma unroll 1
			for (uint32_t s = 0; s < T; s++) {
				uint32_t st = s + K;
				for (uint32_t k = 0; k < w; k++) st = REPC_STEP1(st, __ldg(p + k));
				st = (HAS_DEAD && st == deadK) ? REP_DIED_PREFIX : st - K;
				img |= (uint64_t) st << (4u * s);
			}
		}
		for (uint32_t s = 0; s < T; s++) {
			const uint32_t v = (uint32_t) (img >> (4u * s)) & 15u;
			if (v < REP_DIED_BODY && !((a.absorb_mask >> v) & 1u)) live |= 1u << v;
		}
		if (beg + w >= end) live = 0;
	}

	while (__any_sync(0xFFFFFFFFu, live != 0)) {
		if (tiled) {
			/* every lane has a full chunk of C bytes: nsec sectors after the prefix, the same for all lanes.
			 * Lanes without a (further) live image walk along from the dead row / their last state and
			 * discard the result -- the ring is a warp-wide protocol. */
			const bool mine = live != 0;
			const uint32_t v = mine ? (uint32_t) __ffs((int) live) - 1u : 0u;
			live &= live - 1u;
			uint32_t st = v + K;
			bool died = false;
			uint32_t from = 0;
			uint64_t dpos = 0;
			const uint32_t nsec = (uint32_t) ((a.C - REP_W) >> 5);
			const uint32_t y = gw * 32u;
			/* prologue: NST tiles in flight */
			if (lane == 0) {
#pragma unroll
				for (int s = 0; s < NST; s++) {
					if ((uint32_t) s < nsec) {
						const uint32_t slot = (uses + (uint32_t) s) % (uint32_t) NST;
						mbar_expect_tx(bars_a + 8u * slot, REPC_STAGE_BYTES);
						repc_tma_tile(ring_a + slot * REPC_STAGE_BYTES, &tmap, REP_W + 32u * (uint32_t) s, y, bars_a + 8u * slot);
					}
				}
			}
			for (uint32_t i = 0; i < nsec; i++) {
				const uint32_t slot = uses % (uint32_t) NST, parity = (uses / (uint32_t) NST) & 1u;
				mbar_wait(bars_a + 8u * slot, parity);
				const uint32_t row_a = ring_a + slot * REPC_STAGE_BYTES + lane * 32u;
				const uint4 x0 = lds_u128(row_a + sw16), x1 = lds_u128(row_a + (sw16 ^ 16u));
				const uint32_t entry_ = st;
				REPC_STEP4(st, x0.x); REPC_STEP4(st, x0.y); REPC_STEP4(st, x0.z); REPC_STEP4(st, x0.w);
				REPC_STEP4(st, x1.x); REPC_STEP4(st, x1.y); REPC_STEP4(st, x1.z); REPC_STEP4(st, x1.w);
				/* the walk has consumed all eight words: the slot may be overwritten */
				__syncwarp();
				uses++;
				if (lane == 0 && i + (uint32_t) NST < nsec) {
					mbar_expect_tx(bars_a + 8u * slot, REPC_STAGE_BYTES);
					repc_tma_tile(ring_a + slot * REPC_STAGE_BYTES, &tmap, REP_W + 32u * (i + (uint32_t) NST), y, bars_a + 8u * slot);
				}
				if (HAS_DEAD && st == deadK && !died && mine) {
					/* a byte of this sector had no edge: re-walk it (from global memory) to find which;
					 * the dead row absorbs, so the lane simply keeps walking along afterwards */
					const uint8_t *p = a.buf + beg + REP_W + 32ull * i;
					uint32_t s2 = entry_;
					for (int k = 0; k < 32; k++) {
						const uint32_t nx = REPC_STEP1(s2, __ldg(p + k));
						if (nx == deadK) { died = true; from = s2; dpos = beg + REP_W + 32ull * i + (uint64_t) k; break; }
						s2 = nx;
					}
				}
			}
			if (mine) {
				uint32_t e = st - K;
				if (died) {
					a.body_off[(size_t) c * 16u + v] = dpos;
					a.body_from[(size_t) c * 16u + v] = (uint8_t) (from - K);
					e = REP_DIED_BODY;
				}
				exit_of = (exit_of & ~(15ull << (4u * v))) | ((uint64_t) e << (4u * v));
			}
		} else if (live != 0) {
			const uint32_t v = (uint32_t) __ffs((int) live) - 1u;
			live &= live - 1u;
			uint32_t st = v + K;
			uint64_t pos = beg + w;
			bool died = false;
			uint32_t from = 0;
			while (pos < end && (reinterpret_cast<uintptr_t>(a.buf + pos) & 31u) != 0) {
				const uint32_t nx = REPC_STEP1(st, __ldg(a.buf + pos));
				if (HAS_DEAD && nx == deadK) { died = true; from = st; break; }
				st = nx; pos++;
			}
			if (!died) {
				const uint8_t *p = a.buf + pos;
				const uint32_t nsec = (uint32_t) ((end - pos) >> 5);
				uint32_t A[8], B[8];
				uint32_t i = 0;
				bool stop = false;
#define REPC_WALK(X)                                                                            \
	do {                                                                                        \
		const uint32_t entry_ = st;                                                             \
		_Pragma("unroll") for (int k = 0; k < 8; k++) REPC_STEP4(st, X[k]);                     \
		if (HAS_DEAD && st == deadK) {                                                          \
			st = entry_;                                                                        \
			for (int k = 0; k < 32; k++) {                                                      \
				const uint32_t nx = REPC_STEP1(st, __ldg(p + k));                               \
				if (nx == deadK) { died = true; from = st; p += k; break; }                     \
				st = nx;                                                                        \
			}                                                                                   \
			stop = true;                                                                        \
		} else {                                                                                \
			p += 32; i++;                                                                       \
		}                                                                                       \
	} while (0)
				if (nsec > 0) ld256(p, A);
				while (i < nsec) {
					if (i + 1 < nsec) ld256(p + 32, B);
					REPC_WALK(A);
					if (stop || i >= nsec) break;
					if (i + 1 < nsec) ld256(p + 32, A);
					REPC_WALK(B);
					if (stop) break;
				}
#undef REPC_WALK
				pos = (uint64_t) (p - a.buf);
			}
			if (!died) {
				for (; pos < end; pos++) {
					const uint32_t nx = REPC_STEP1(st, __ldg(a.buf + pos));
					if (HAS_DEAD && nx == deadK) { died = true; from = st; break; }
					st = nx;
				}
			}
			uint32_t e = st - K;
			if (died) {
				a.body_off[(size_t) c * 16u + v] = pos;
				a.body_from[(size_t) c * 16u + v] = (uint8_t) (from - K);
				e = REP_DIED_BODY;
			}
			exit_of = (exit_of & ~(15ull << (4u * v))) | ((uint64_t) e << (4u * v));
		}
	}

	uint64_t cmap = 0xFEDCBA9876543210ull;
	if (have) {
		cmap = 0;
		for (uint32_t s = 0; s < T; s++) {
			const uint32_t v = (uint32_t) (img >> (4u * s)) & 15u;
			const uint32_t e = v == REP_DIED_PREFIX ? REP_DIED_PREFIX : (uint32_t) (exit_of >> (4u * v)) & 15u;
			cmap |= (uint64_t) e << (4u * s);
		}
	}
	wm[warp * 32u + lane] = cmap;
	__syncwarp();
	if (lane < T) {
		uint32_t st = lane, dc = 0xFFFFFFFFu, ds = 0;
		for (uint32_t l = 0; l < 32; l++) {
			const uint32_t e = (uint32_t) (wm[warp * 32u + l] >> (4u * st)) & 15u;
			if (e >= REP_DIED_BODY) { dc = gw * 32u + l; ds = st; st = 0xFFu; break; }
			st = e;
		}
		s_wmap[warp * 16u + lane] = (uint8_t) st;
		s_wdc[warp * 16u + lane] = dc;
		s_wds[warp * 16u + lane] = (uint8_t) ds;
	}
	__syncwarp();
	uint32_t last = 0;
	if (lane == 0) {
		__threadfence_block();
		last = atomicAdd(s_done, 1u) == nactive - 1u ? 1u : 0u;
		__threadfence_block();
	}
	last = __shfl_sync(0xFFFFFFFFu, last, 0);
	if (last && lane < T) {
		uint32_t st = lane, dc = 0xFFFFFFFFu, ds = 0;
		for (uint32_t wv = 0; wv < nactive; wv++) {
			const uint32_t e = *reinterpret_cast<volatile uint8_t *>(s_wmap + wv * 16u + st);
			if (e == 0xFFu) {
				dc = *reinterpret_cast<volatile uint32_t *>(s_wdc + wv * 16u + st);
				ds = *reinterpret_cast<volatile uint8_t *>(s_wds + wv * 16u + st);
				st = 0xFFu;
				break;
			}
			st = e;
		}
		a.wmap[(size_t) blockIdx.x * 16u + lane] = (uint8_t) st;
		a.wdc[(size_t) blockIdx.x * 16u + lane] = dc;
		a.wds[(size_t) blockIdx.x * 16u + lane] = (uint8_t) ds;
	}
}

#undef REPC_STEP4
#undef REPC_STEP1

} // namespace fsmb200
#endif
