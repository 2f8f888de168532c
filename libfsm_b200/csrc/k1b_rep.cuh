/*
 * k1b_rep.cuh -- K1b for SMALL automata (at most 12 table rows): the whole stream map in one pass.
 *
 * Same monoid formulation as k1b_stream.cu (src/libfsm/exec.c:132-151 is one serial chain; a byte
 * range maps entry states to exit states and maps compose), but fused: ONE kernel walks the chunks and
 * folds the chunk maps of a warp, one small kernel folds the warp maps.  No job lists, no memsets.
 *
 * Why a separate form: the generic body (k1_lane_kernel over the jobs) keeps ONE copy of the table in
 * shared memory, and the lanes of a warp -- each in its own state, on its own byte -- collide on its
 * banks: 1.84 wavefronts per lookup on the UTF-8 validator, L1TEX 98.8 % busy (profiles/
 * r1_k1b_body_utf8_ncu_full.txt).  A table this small can be REPLICATED PER LANE so that lane l only
 * ever touches bank l: exactly one wavefront per lookup, whatever the states and bytes.
 *
 *   address of entry (state, byte) of lane l  =  state << 14 | (byte >> 2) << 8 | l << 2 | (byte & 3)
 *
 * Bits 2..6 are the bank and depend on the lane alone.  Bit 7 is left unused on purpose: it puts the
 * word index of the byte (byte >> 2) on a byte boundary of the address, so that ONE prmt.b32 per input
 * byte builds the whole byte-dependent part of the address {byte & 3 | lane bits, byte >> 2, 0, 0} from
 * two per-word registers -- 3.75 instructions per input byte (3 per word + PRMT, LEA, LDS.U8 per byte),
 * the dependent chain being LEA + LDS.  16 KiB per table row: 9 rows (UTF-8 validator + dead row) =
 * 144 KiB of the 227 KiB of shared memory, at most 12 rows.
 *
 * Per lane (one chunk of C bytes; C = stream length / lanes of the grid, so the grid is ONE full wave):
 *   prefix  the first 64 bytes from every entry state (chains from wrong states die or merge), out of
 *           two 256-bit loads held in registers;
 *   body    the rest of the chunk from each DISTINCT live, non-absorbing image (one for practical
 *           automata; more only cost time, never exactness), one aligned 32-byte sector per 256-bit load
 *           into 2..4 ROTATING register buffers -- with one load in flight per lane 31 % of the warp time
 *           sat on the first use of a loaded register; four buffers (three loads in flight) is the default:
 *           2 GiB in 0.60 / 0.56 / 0.50 ms with 2 / 3 / 4.  A walk that meets the dead row finds the exact
 *           offset in the sector it died in and records it;
 *   fold    the chunk map (4 bits per entry state) goes to shared memory and lanes 0..T-1 fold the 32
 *           maps of the warp in order; the last warp of a CTA to finish folds the CTA's warp maps.
 * k1b_rep_final_kernel folds the CTA maps (fan-in 32 per level, in shared memory) and resolves a death
 * on the true path to its stream offset (re-walking at most one 64-byte prefix).
 * Measured (profiles/r2_k1b_rep_*): 2 GiB of UTF-8 through the 8-state validator in 0.50 ms = 4.29 TB/s =
 * 0.65 of the measured HBM peak (generic path 0.935 ms); then L1TEX 85 %: 43 points table lookups (one
 * wavefront each), 28 points the 256-bit loads (21 data-pipe wavefronts per request of 32 different lines).
 */
#ifndef FSM_B200_K1B_REP_CUH
#define FSM_B200_K1B_REP_CUH

#include "common.h"
#include "k1_device.cuh"

namespace fsmb200 {

constexpr uint32_t REP_MAX_ROWS = 12;        /* 12 x 16 KiB of table + 8 KiB of chunk maps */
constexpr uint32_t REP_W = 64;               /* prefix window */
constexpr uint32_t REP_ROW_SHIFT = 14;
constexpr uint32_t REP_CMAPS_BYTES = 32u * 32u * 8u;  /* chunk maps of a CTA (4 bits per entry state) */
constexpr uint32_t REP_MAPS_BYTES = REP_CMAPS_BYTES + 32u * 16u * 6u + 256u;   /* + warp maps with death records + a counter: in front of the table */
constexpr uint32_t REP_DIED_BODY = 0xE, REP_DIED_PREFIX = 0xF;

struct StreamOut {          /* per entry state, [T] */
	uint32_t state;         /* exit state, or dead-from state when died */
	uint32_t died;
	uint64_t dead_off;      /* offset within the range of the first byte without an edge */
};

struct RepArgs {
	const uint8_t *buf;
	uint64_t len, C;
	uint32_t mis;            /* address of buf modulo 32: chunk c > 0 starts at c * C - mis (sector-aligned) */
	uint32_t nchunks, nwarps;
	uint32_t wpc, nmaps;     /* warps per CTA in use; CTAs = maps the final kernel folds */
	uint32_t T, ntable, dead;   /* dead = ntable - 1 for incomplete automata, else NO_EDGE */
	uint32_t absorb_mask;    /* bit s: every byte loops state s back to itself */
	const uint8_t *dense;    /* [ntable][256] next-state bytes (dead row included) */
	/* CTA maps (each CTA folds the maps of its warps) */
	uint8_t *wmap;           /* [nmaps][16]: exit state per entry state, 0xFF = died */
	uint32_t *wdc;           /* [nmaps][16]: chunk in which it died */
	uint8_t *wds;            /* [nmaps][16]: state in which that chunk was entered */
	/* deaths in a body walk, by (chunk, image state); written only when it happens */
	uint64_t *body_off;      /* [nchunks][16] stream offset of the byte without an edge */
	uint8_t *body_from;      /* [nchunks][16] state it was read in */
	/* final fold */
	uint32_t *lv_dc[2];      /* [ceil(nmaps/32)][16] ping-pong death records of the upper levels */
	uint8_t *lv_ds[2];
	StreamOut *out;          /* [T] */
};

__device__ __forceinline__ uint32_t
prmt_sx(uint32_t a, uint32_t b, uint32_t sel)
{
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));   /* selector bit 3 = sign fill */
	return d;
}

__device__ __forceinline__ uint64_t
rep_chunk_beg(const RepArgs &a, uint32_t c)
{
	return c == 0 ? 0ull : (uint64_t) c * a.C - a.mis;
}

__device__ __forceinline__ uint64_t
rep_chunk_end(const RepArgs &a, uint32_t c)
{
	const uint64_t e = (uint64_t) (c + 1) * a.C - a.mis;
	return e < a.len ? e : a.len;
}

/* The table starts at a 16 KiB-aligned SHARED address and its entries hold (next state + table address
 * >> 14): the state register is the row's address bits, and the dependent chain is exactly
 * LEA (state << 14 + byte part) -> LDS.U8 -- no base add. */
__device__ __forceinline__ uint32_t
lds_u8(uint32_t addr)
{
	uint32_t v;
	asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}

/* one byte from a global pointer (heads, tails, the sector a walk died in) */
#define REP_STEP1(st, b) lds_u8(((st) << REP_ROW_SHIFT) + (((uint32_t) (b) >> 2) << 8) + ((uint32_t) (b) & 3u) + lane4)

/* four bytes of a word: per word SHF + LOP (word indices) and one LOP3 (byte selects | lane bits), per
 * byte PRMT (address bytes {sel | lane, index, 0, 0}: the index bytes are below 0x40, so their sign
 * fill is the zero we need) + LEA + LDS.U8 */
#define REP_STEP4(st, w)                                                  \
	do {                                                                  \
		const uint32_t h_ = ((w) >> 2) & 0x3F3F3F3Fu;                     \
		const uint32_t q_ = ((w) & 0x03030303u) | lanebits;               \
		st = lds_u8(((st) << REP_ROW_SHIFT) + prmt_sx(q_, h_, 0xCC40u));  \
		st = lds_u8(((st) << REP_ROW_SHIFT) + prmt_sx(q_, h_, 0xCC51u));  \
		st = lds_u8(((st) << REP_ROW_SHIFT) + prmt_sx(q_, h_, 0xCC62u));  \
		st = lds_u8(((st) << REP_ROW_SHIFT) + prmt_sx(q_, h_, 0xCC73u));  \
	} while (0)

/* 256-bit sector load; HINT: ask L2 to fetch 256 B around a miss (the lane reads the next seven sectors of
 * that run next, so they become L2 hits instead of DRAM round trips) */
template <int HINT>
__device__ __forceinline__ void
rep_ld256(const uint8_t *p, uint32_t (&w)[8])
{
	if (HINT) {
		asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		    : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
	} else {
		ld256(p, w);
	}
}

template <bool HAS_DEAD, int HINT, int NBUF>
__global__ void __launch_bounds__(1024, 1)
k1b_rep_kernel(const RepArgs a)
{
	extern __shared__ __align__(1024) uint8_t dsm[];
	uint64_t *wm = reinterpret_cast<uint64_t *>(dsm);        /* [32 warps][32 lanes] chunk maps, 8 KiB */
	uint32_t *s_wdc = reinterpret_cast<uint32_t *>(dsm + REP_CMAPS_BYTES);            /* [32 warps][16] */
	uint8_t *s_wmap = dsm + REP_CMAPS_BYTES + 32u * 16u * 4u;                         /* [32 warps][16] */
	uint8_t *s_wds = s_wmap + 32u * 16u;                                              /* [32 warps][16] */
	uint32_t *s_done = reinterpret_cast<uint32_t *>(s_wds + 32u * 16u);               /* warps of this CTA that have folded */
	if (threadIdx.x == 0) *s_done = 0;
	const uint32_t rep_base = (smem_u32(dsm) + REP_MAPS_BYTES + 16383u) & ~16383u;   /* shared address of the table */
	const uint32_t K = rep_base >> REP_ROW_SHIFT;            /* states live in registers as state + K */

	/* replicate the table: word j of row s of lane l at s << 14 | j << 8 | l << 2.  Consecutive threads
	 * take consecutive lanes: the global read is a broadcast, the shared store conflict-free. */
	{
		const uint32_t *d32 = reinterpret_cast<const uint32_t *>(a.dense);
		const uint32_t total = a.ntable * 64u * 32u;
		const uint32_t k4 = K * 0x01010101u;
		for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
			const uint32_t l = i & 31u, j = (i >> 5) & 63u, s = i >> 11;
			const uint32_t v = __ldg(d32 + s * 64u + j) + k4;       /* bytes stay below 256: at most 12 + K */
			asm volatile("st.shared.u32 [%0], %1;" :: "r"(rep_base + ((s << REP_ROW_SHIFT) | (j << 8) | (l << 2))), "r"(v) : "memory");
		}
	}
	__syncthreads();

	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	/* CTA b owns the consecutive warp maps [b * wpc, b * wpc + wpc): it folds them itself (below) */
	const uint32_t gw = blockIdx.x * a.wpc + warp;
	if (warp >= a.wpc || gw >= a.nwarps) return;
	const uint32_t nactive = min(a.wpc, a.nwarps - blockIdx.x * a.wpc);
	const uint32_t lane4 = lane << 2;
	const uint32_t lanebits = lane4 * 0x01010101u;
	const uint32_t T = a.T;
	const uint32_t deadK = a.dead + K;           /* NO_EDGE + K never equals a state */
	const uint32_t c = gw * 32u + lane;
	const bool have = c < a.nchunks;
	const uint64_t beg = have ? rep_chunk_beg(a, c) : 0, end = have ? rep_chunk_end(a, c) : 0;

	uint64_t img = 0;            /* 4 bits per entry state: state after the prefix, or REP_DIED_PREFIX */
	uint64_t exit_of = 0xFEDCBA9876543210ull;   /* 4 bits per image state: where the body walk from it ends */
	uint32_t live = 0;           /* image states that need a body walk */
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
				for (int k = 0; k < 8; k++) REP_STEP4(st, A[k]);
#pragma unroll
				for (int k = 0; k < 8; k++) REP_STEP4(st, B[k]);
				st = (HAS_DEAD && st == deadK) ? REP_DIED_PREFIX : st - K;
				img |= (uint64_t) st << (4u * s);
			}
		} else {                 /* the first chunk of an unaligned buffer, a short last chunk */
#pragma unroll 1
			for (uint32_t s = 0; s < T; s++) {
				uint32_t st = s + K;
				for (uint32_t k = 0; k < w; k++) st = REP_STEP1(st, __ldg(p + k));
				st = (HAS_DEAD && st == deadK) ? REP_DIED_PREFIX : st - K;
				img |= (uint64_t) st << (4u * s);
			}
		}
		for (uint32_t s = 0; s < T; s++) {
			const uint32_t v = (uint32_t) (img >> (4u * s)) & 15u;
			if (v < REP_DIED_BODY && !((a.absorb_mask >> v) & 1u)) live |= 1u << v;
		}
		if (beg + w >= end) live = 0;        /* the prefix was the whole chunk */
	}

	/* body: one walk per distinct live image (the j-th of every lane at the same time) */
	while (__any_sync(0xFFFFFFFFu, live != 0)) {
		if (live != 0) {
			const uint32_t v = (uint32_t) __ffs((int) live) - 1u;
			live &= live - 1u;
			uint32_t st = v + K;
			uint64_t pos = beg + w;
			bool died = false;
			uint32_t from = 0;
			/* head: up to the first sector boundary (only the first chunk of an unaligned buffer) */
			while (pos < end && (reinterpret_cast<uintptr_t>(a.buf + pos) & 31u) != 0) {
				const uint32_t nx = REP_STEP1(st, __ldg(a.buf + pos));
				if (HAS_DEAD && nx == deadK) { died = true; from = st; break; }
				st = nx; pos++;
			}
			if (!died) {
				/* sectors: buffers that swap roles (no register copies), loads issued ahead of the walk */
				const uint8_t *p = a.buf + pos;
				const uint32_t nsec = (uint32_t) ((end - pos) >> 5);       /* a chunk is shorter than 4 GiB */
				uint32_t A[8], B[8];
				uint32_t i = 0;
				bool stop = false;
#define REP_WALK(X)                                                                             \
	do {                                                                                        \
		const uint32_t entry_ = st;                                                             \
		_Pragma("unroll") for (int k = 0; k < 8; k++) REP_STEP4(st, X[k]);                      \
		if (HAS_DEAD && st == deadK) {                                                          \
			/* a byte of this sector had no edge: re-walk it to find which */                   \
			st = entry_;                                                                        \
			for (int k = 0; k < 32; k++) {                                                      \
				const uint32_t nx = REP_STEP1(st, __ldg(p + k));                                \
				if (nx == deadK) { died = true; from = st; p += k; break; }                     \
				st = nx;                                                                        \
			}                                                                                   \
			stop = true;                                                                        \
		} else {                                                                                \
			p += 32; i++;                                                                       \
		}                                                                                       \
	} while (0)
				if (NBUF == 4) {
					uint32_t C[8], D[8];
					if (nsec > 0) rep_ld256<HINT>(p, A);
					if (nsec > 1) rep_ld256<HINT>(p + 32, B);
					if (nsec > 2) rep_ld256<HINT>(p + 64, C);
					while (i < nsec) {
						if (i + 3 < nsec) rep_ld256<HINT>(p + 96, D);
						REP_WALK(A);
						if (stop || i >= nsec) break;
						if (i + 3 < nsec) rep_ld256<HINT>(p + 96, A);
						REP_WALK(B);
						if (stop || i >= nsec) break;
						if (i + 3 < nsec) rep_ld256<HINT>(p + 96, B);
						REP_WALK(C);
						if (stop || i >= nsec) break;
						if (i + 3 < nsec) rep_ld256<HINT>(p + 96, C);
						REP_WALK(D);
						if (stop) break;
					}
				} else if (NBUF == 3) {
					/* three buffers: the load of sector i + 2 is issued before sector i is walked -- two
					 * sectors per lane in flight (one is not enough to cover the loaded DRAM latency:
					 * 31 % of the warp time sat on the first use of a loaded register) */
					uint32_t C[8];
					if (nsec > 0) rep_ld256<HINT>(p, A);
					if (nsec > 1) rep_ld256<HINT>(p + 32, B);
					while (i < nsec) {
						if (i + 2 < nsec) rep_ld256<HINT>(p + 64, C);
						REP_WALK(A);
						if (stop || i >= nsec) break;
						if (i + 2 < nsec) rep_ld256<HINT>(p + 64, A);
						REP_WALK(B);
						if (stop || i >= nsec) break;
						if (i + 2 < nsec) rep_ld256<HINT>(p + 64, B);
						REP_WALK(C);
						if (stop) break;
					}
				} else {
					if (nsec > 0) rep_ld256<HINT>(p, A);
					while (i < nsec) {
						if (i + 1 < nsec) rep_ld256<HINT>(p + 32, B);
						REP_WALK(A);
						if (stop || i >= nsec) break;
						if (i + 1 < nsec) rep_ld256<HINT>(p + 32, A);
						REP_WALK(B);
						if (stop) break;
					}
				}
#undef REP_WALK
				pos = (uint64_t) (p - a.buf);
			}
			if (!died) {
				for (; pos < end; pos++) {
					const uint32_t nx = REP_STEP1(st, __ldg(a.buf + pos));
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

	/* the chunk map: entry state -> exit state / died */
	uint64_t cmap = 0xFEDCBA9876543210ull;       /* a chunk beyond the end: identity */
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
	/* the last warp of the CTA to get here folds the CTA's warp maps in order (no barrier: warps without
	 * chunks have left, and nobody waits for the slowest walk) */
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

#undef REP_STEP4
#undef REP_STEP1

/* Fold the CTA maps in order (fan-in 32 per level, maps in shared memory), then turn a death on the path
 * of entry state s into (offset, state): the chunk it happened in is re-walked over its prefix; if the walk
 * survives that, the body walk from the image recorded where it died. */
__global__ void __launch_bounds__(1024, 1)
k1b_rep_final_kernel(const RepArgs a)
{
	extern __shared__ __align__(16) uint8_t fm[];
	const uint32_t T = a.T;
	uint32_t n = a.nmaps;
	uint8_t *A = fm, *B = fm + (size_t) a.nmaps * 16u;
	for (uint32_t i = threadIdx.x; i < n * 4u; i += blockDim.x) {
		reinterpret_cast<uint32_t *>(A)[i] = reinterpret_cast<const uint32_t *>(a.wmap)[i];
	}
	__syncthreads();
	const uint32_t *dcA = a.wdc;
	const uint8_t *dsA = a.wds;
	int pp = 0;
	while (n > 1) {
		const uint32_t n2 = (n + 31u) / 32u;
		uint32_t *dcB = a.lv_dc[pp];
		uint8_t *dsB = a.lv_ds[pp];
		for (uint32_t idx = threadIdx.x; idx < n2 * T; idx += blockDim.x) {
			const uint32_t g = idx / T, s = idx % T;
			const uint32_t hi = min(n, g * 32u + 32u);
			uint32_t st = s;
			for (uint32_t i = g * 32u; i < hi; i++) {
				const uint32_t e = A[i * 16u + st];
				if (e == 0xFFu) {
					dcB[g * 16u + s] = dcA[i * 16u + st];
					dsB[g * 16u + s] = dsA[i * 16u + st];
					st = 0xFFu;
					break;
				}
				st = e;
			}
			B[g * 16u + s] = (uint8_t) st;
		}
		__syncthreads();          /* one CTA: the global death records written above are visible after the barrier */
		uint8_t *t = A; A = B; B = t;
		dcA = dcB; dsA = dsB;
		pp ^= 1;
		n = n2;
	}
	const uint32_t s = threadIdx.x;
	if (s >= T) return;
	const uint32_t e = A[s];
	if (e != 0xFFu) {
		a.out[s].state = e; a.out[s].died = 0; a.out[s].dead_off = 0xFFFFFFFFFFFFFFFFull;
		return;
	}
	const uint32_t c = dcA[s];
	uint32_t st = dsA[s];
	const uint64_t beg = rep_chunk_beg(a, c), end = rep_chunk_end(a, c);
	const uint32_t w = (uint32_t) min((uint64_t) REP_W, end - beg);
	for (uint32_t k = 0; k < w; k++) {
		const uint32_t nx = a.dense[st * 256u + a.buf[beg + k]];
		if (nx == a.dead) {
			a.out[s].state = st; a.out[s].died = 1; a.out[s].dead_off = beg + k;
			return;
		}
		st = nx;
	}
	a.out[s].state = a.body_from[(size_t) c * 16u + st];
	a.out[s].died = 1;
	a.out[s].dead_off = a.body_off[(size_t) c * 16u + st];
}

} // namespace fsmb200
#endif
