/*
 * k1_lines.cu -- K1 for RAGGED batches (many short inputs at arbitrary alignment: the log lines of
 * BASELINE config 3) and for EAGER OUTPUTS, on tables that fit shared memory.
 *
 * One reference fsm_exec call per line (src/libfsm/exec.c:85-167); with eager outputs the ids of
 * the start state and of every state entered fire as well (exec.c:55-83,126-130,140-144;
 * include/fsm/fsm.h:273-336) and the answer carries their SET as a bitset per line.
 *
 * Lane l of a warp walks lines l, l+32, ... of the warp's contiguous range, ONE aligned 32-byte
 * sector per loop iteration (256-bit load = one full DRAM sector), three sector buffers: A (being
 * walked), B (next sector of the line), N (first sector of the lane's next line).
 *
 * What keeps the per-byte cost at PRMT + LDS + IMAD + LDS (+ VIMNMX):
 *   - rows are indexed by byte class and have one extra NOP column in which every state loops to
 *     itself.  Bytes of a sector that lie outside the line are not skipped: their LUT index gets
 *     256 added (PRMT pulls the high byte from a per-word validity word), which maps to NOP -- no
 *     select on the dependent state chain, no byte-wise head / tail path;
 *   - dfa_compile.cu numbers the states with eager outputs last, just before the dead row.  Per
 *     byte the kernel keeps max(st) and the sum of max(st, first_event - 1): after the sector,
 *       no eager outputs (EV_DEAD): the dead row absorbs, so the sum counts the steps spent dead and
 *         gives the offset of the missing edge; the state it was taken from is 0..3 exact steps away
 *         from a per-word snapshot of the walk;
 *       eager outputs (EV_EAGER): sum == max - (first_event - 1) <=> exactly one state with ids was
 *         entered, once (86 % of the sectors of BASELINE config 3 enter none, 13.8 % one, 0.5 % more):
 *         its ids are OR-ed in from a small global array; anything else re-walks the sector byte by
 *         byte, out of line;
 *   - every lane asks L2 for the cache line LINES_PREFETCH bytes ahead of the sector it reads
 *     (prefetch.global.L2): the lanes of a warp sweep a contiguous region of the batch, so together
 *     they prefetch the region the warp reads next, and the 256-bit loads of the walk hit L2.
 *
 * Algorithmic bytes per line: its bytes, read once, + 16 B record (+ 8 W B id bitset).  Bound:
 * the shared-memory lookup rate (two dependent-free + one dependent LDS per byte), see DESIGN.md.
 */
#include <cstring>

#include "common.h"
#include "k1_exec_batch.h"
#include "k1_device.cuh"

using namespace fsmb200;

namespace {

struct LinesArgs {
	const uint8_t *blob;          /* [512 B LUT][rows][is_end] */
	uint32_t blob_bytes, pitch /* bytes */, end_off, first_event, dead, start;
	const uint32_t *perm_inv;     /* new state number -> caller's */
	const uint8_t *absorb;        /* by new number, or nullptr */
	const uint64_t *masks;        /* [ntable][W] by the caller's numbering (global memory) */
	const uint64_t *ev_masks;     /* [ntable - first_event][W] by new number - first_event */
	uint32_t prefetch;            /* L2 prefetch distance in bytes, 0 = off */
	uint64_t start_mask[4];
	const uint8_t *base;
	const uint64_t *offsets;      /* n + 1 entries, or nullptr: fixed stride */
	uint64_t stride, len, n;
	fsm_b200_result *out;
	uint64_t *out_masks;          /* [n][W] */
};

/* the staged blob; file scope so that the out-of-line re-walk addresses it as shared memory too */
extern __shared__ __align__(1024) uint8_t lines_smem[];

/* {x.byte t, v.byte t, 0, 0}: PRMT with the sign-fill mode for the two upper bytes (v.byte t is 0 or 1) */
template <int T>
__device__ __forceinline__ uint32_t
byte_and_flag(uint32_t x, uint32_t v)
{
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(v), "n"(T | ((4 + T) << 4) | ((0xC + T) << 8) | ((0xC + T) << 12)));
	return d;
}

template <int W> struct Rewalk {
	uint32_t st, at, died;
	uint64_t m[W > 0 ? W : 1];
};

/* Exact walk of one sector: first byte (within `mask`) without an edge and the state it was taken
 * from, ids of every state entered.  Cold path. */
template <typename E, int W>
__device__ __noinline__ Rewalk<W>
lines_rewalk(uint32_t pitch, uint32_t entry, uint32_t mask,
	uint32_t first_event, uint32_t dead, const uint32_t *perm_inv, const uint64_t *masks, const uint32_t (&w)[8])
{
	const uint8_t *lut = lines_smem;
	Rewalk<W> r;
	r.st = entry; r.at = 32; r.died = 0;
#pragma unroll
	for (int k = 0; k < (W > 0 ? W : 1); k++) r.m[k] = 0;
#pragma unroll 1
	for (uint32_t j = 0; j < 32; j++) {
		if (!((mask >> j) & 1u)) continue;
		const uint32_t b = (w[j >> 2] >> (8u * (j & 3u))) & 0xFFu;
		const uint32_t nx = (uint32_t) *reinterpret_cast<const E *>(lines_smem + 512 + r.st * pitch + lut[b]);
		if (nx == dead) { r.died = 1; r.at = j; break; }
		r.st = nx;
		if (W > 0 && nx >= first_event) {
			const uint32_t old = __ldg(perm_inv + nx);
#pragma unroll
			for (int k = 0; k < W; k++) r.m[k] |= __ldg(masks + (size_t) old * W + k);
		}
	}
	return r;
}

constexpr int LINES_THREADS = 768;

enum { EV_NONE = 0, EV_DEAD = 1, EV_EAGER = 2 };

template <typename E, int W, int EV>
__global__ void __launch_bounds__(LINES_THREADS, 1)
k1_lines_kernel(const LinesArgs a)
{
	__shared__ uint64_t blob_bar;
	stage_blob(lines_smem, a.blob, a.blob_bytes, &blob_bar);
	const uint8_t *lut = lines_smem;                  /* byte (+256 when outside the line) -> class * sizeof(E) */
	const uint8_t *tab = lines_smem + 512;
	const uint8_t *is_end = lines_smem + a.end_off;
	const uint32_t pitch = a.pitch;                   /* bytes */

	const uint32_t lane = threadIdx.x & 31u;
	const uint64_t nwarps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
	const uint64_t gw = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	/* contiguous range of lines per warp, in multiples of 32 */
	const uint64_t per = (((a.n + nwarps - 1) / nwarps) + 31u) & ~31ull;
	const uint64_t wbeg = gw * per;
	const uint64_t wend = min(a.n, wbeg + per);
	const uintptr_t base = reinterpret_cast<uintptr_t>(a.base);
	/* the bytes this batch may touch: sectors reaching outside are assembled from byte loads */
	const uintptr_t lo_ptr = base + (a.offsets != nullptr ? a.offsets[0] : 0);
	const uintptr_t hi_ptr = base + (a.offsets != nullptr ? a.offsets[a.n] : (a.n - 1) * a.stride + a.len);

	auto bounds = [&](uint64_t i, uintptr_t &b, uintptr_t &e) {
		if (a.offsets != nullptr) { b = base + a.offsets[i]; e = base + a.offsets[i + 1]; }
		else { b = base + i * a.stride; e = b + a.len; }
	};
	auto load_sector = [&](uintptr_t saddr, uint32_t (&w)[8]) {
		if (saddr >= lo_ptr && saddr + 32 <= hi_ptr) {
			ld256(reinterpret_cast<const uint8_t *>(saddr), w);
		} else {
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
		bounds(i, cur, end);
		load_sector(cur & ~(uintptr_t) 31, A);
		have_next = i + 32 < wend;
		if (have_next) {
			bounds(i + 32, nbeg, nend);
			load_sector(nbeg & ~(uintptr_t) 31, N);
		}
	}
	uint32_t st = a.start;
	uintptr_t line_beg = cur;
	uint64_t acc[W > 0 ? W : 1];
#pragma unroll
	for (int k = 0; k < (W > 0 ? W : 1); k++) acc[k] = W > 0 ? a.start_mask[k] : 0;

	while (have) {
		const uintptr_t saddr = cur & ~(uintptr_t) 31;
		const uint32_t lo = (uint32_t) (cur - saddr);
		const uint32_t hi = (end - saddr) >= 32 ? 32u : (uint32_t) (end - saddr);
		const bool more = saddr + 32 < end;              /* the line continues in the next sector */
		if (more) load_sector(saddr + 32, B);

		if (a.prefetch != 0 && (saddr & 96u) == 0 && saddr + a.prefetch + 128 <= hi_ptr) {
			asm volatile("prefetch.global.L2 [%0];" :: "l"(saddr + a.prefetch));
		}

		/* walk all 32 bytes; bytes outside [lo, hi) take the NOP column */
		const uint32_t mask = (hi > lo) ? ((0xFFFFFFFFu << lo) & (0xFFFFFFFFu >> (32u - hi))) : 0u;
		const uint32_t inv = ~mask;
		const uint32_t entry = st;
		const uint32_t base = a.first_event - 1u;         /* EV_DEAD: first_event is the dead row */
		uint32_t seen = 0, ssum = 0;
		uint32_t snap[8];
#pragma unroll
		for (int k = 0; k < 8; k++) {
			/* byte t of vw = 1 when byte 4k + t is outside the line */
			const uint32_t vw = (((inv >> (4 * k)) & 0xFu) * 0x00204081u) & 0x01010101u;
			/* index = byte | outside << 8 */
#define LINES_STEP(T)                                                                             \
			st = (uint32_t) *reinterpret_cast<const E *>(tab + st * pitch + lut[byte_and_flag<T>(A[k], vw)]); \
			if (EV == EV_EAGER) seen = max(seen, st);                                             \
			if (EV != EV_NONE) ssum += max(st, base);
			LINES_STEP(0) LINES_STEP(1) LINES_STEP(2) LINES_STEP(3)
#undef LINES_STEP
			if (EV == EV_DEAD) snap[k] = st;
		}
		bool died = false;
		uint32_t consumed_here = hi - lo;
		if (EV == EV_DEAD && st == a.dead) {
			/* the dead row absorbs (NOP steps included): ssum - 32 base = steps spent dead */
			const uint32_t j = 32u - (ssum - 32u * base);          /* sector byte that had no edge */
			const uint32_t kq = j >> 2;
			uint32_t from = entry, word = A[0];
#pragma unroll
			for (int k = 1; k < 8; k++) { if (kq == (uint32_t) k) { from = snap[k - 1]; word = A[k]; } }
#pragma unroll
			for (int t = 0; t < 3; t++) {
				if ((uint32_t) t < (j & 3u) && ((mask >> (4u * kq + (uint32_t) t)) & 1u)) {
					from = (uint32_t) *reinterpret_cast<const E *>(tab + from * pitch + lut[(word >> (8 * t)) & 0xFFu]);
				}
			}
			st = from;
			died = true;
			consumed_here = j - lo;
		}
		if (EV == EV_EAGER && seen >= a.first_event) {
			if (ssum - 32u * base == seen - base && seen != a.dead) {
				/* exactly one state with ids entered, once */
#pragma unroll
				for (int k = 0; k < W; k++) acc[k] |= __ldg(a.ev_masks + (size_t) (seen - a.first_event) * W + k);
			} else {
				const Rewalk<W> r = lines_rewalk<E, W>(pitch, entry, mask, a.first_event, a.dead, a.perm_inv, a.masks, A);
				st = r.st;
#pragma unroll
				for (int k = 0; k < W; k++) acc[k] |= r.m[k];
				if (r.died) { died = true; consumed_here = r.at - lo; }
			}
		}
		/* A state whose 256 edges all loop back to itself keeps the walk where it is whatever
		 * follows: the rest of the line cannot change the record (nor fire a new id), so it is
		 * neither walked nor read. */
		bool absorbed = false;
		if (!died && more && a.absorb != nullptr && __ldg(a.absorb + st)) {
			absorbed = true;
			cur = end;
		} else {
			cur += consumed_here;
		}
		if (died || !more || absorbed) {
			/* line done */
			uint4 v;
			v.x = (!died && is_end[st]) ? 1u : 0u;
			v.y = __ldg(a.perm_inv + st);
			const uint64_t consumed = (uint64_t) (cur - line_beg);
			v.z = (uint32_t) consumed;
			v.w = (uint32_t) (consumed >> 32);
			*reinterpret_cast<uint4 *>(a.out + i) = v;
			if (W > 0) {
#pragma unroll
				for (int k = 0; k < W; k++) { a.out_masks[i * W + k] = acc[k]; acc[k] = a.start_mask[k]; }
			}
			i += 32;
			have = have_next;
			if (have) {
				cur = nbeg; end = nend; line_beg = cur;
				st = a.start;
#pragma unroll
				for (int k = 0; k < 8; k++) A[k] = N[k];
				have_next = i + 32 < wend;
				if (have_next) {
					bounds(i + 32, nbeg, nend);
					load_sector(nbeg & ~(uintptr_t) 31, N);
				}
			}
		} else {
			cur = saddr + 32;
#pragma unroll
			for (int k = 0; k < 8; k++) A[k] = B[k];
		}
	}
}

template <typename E, int W, int EV>
int
launch_lines(const LinesArgs &a, int device, cudaStream_t stream)
{
	auto kern = k1_lines_kernel<E, W, EV>;
	const size_t smem_bytes = (a.blob_bytes + 127u) & ~(size_t) 127u;
	if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_bytes) != cudaSuccess) {
		cudaGetLastError();
		set_error("k1_lines: cannot opt in to %zu bytes of shared memory", smem_bytes);
		errno = EIO;
		return -1;
	}
	int sms = 0;
	FSMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device), return -1);
	int block = LINES_THREADS;
	if (const char *e = getenv("FSM_B200_LINES_BLOCK")) {          /* tuning knob */
		const int v = atoi(e);
		if (v >= 32 && v <= LINES_THREADS && (v % 32) == 0) block = v;
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

template <typename E>
int
dispatch_lines(const LinesArgs &a, uint32_t words, bool dead, int device, cudaStream_t stream)
{
	switch (words) {
	case 0: return dead ? launch_lines<E, 0, EV_DEAD>(a, device, stream) : launch_lines<E, 0, EV_NONE>(a, device, stream);
	case 1: return launch_lines<E, 1, EV_EAGER>(a, device, stream);
	case 2: return launch_lines<E, 2, EV_EAGER>(a, device, stream);
	case 3: return launch_lines<E, 3, EV_EAGER>(a, device, stream);
	case 4: return launch_lines<E, 4, EV_EAGER>(a, device, stream);
	default:
		set_error("k1_lines: %u mask words not supported", words);
		errno = ENOTSUP;
		return -1;
	}
}

} // namespace

namespace fsmb200 {

bool
k1_lines_eligible(const fsm_b200_dfa *dfa)
{
	return dfa->d_lblob != nullptr && getenv("FSM_B200_NO_LINES_KERNEL") == nullptr;
}

/* d_masks == nullptr: plain records only (eager ids, if the DFA has any, are not reported). */
int
k1_lines_launch(const fsm_b200_dfa *dfa, const uint8_t *d_base, const uint64_t *d_offsets,
	uint64_t stride, uint64_t len, size_t n, fsm_b200_result *d_out, uint64_t *d_masks, cudaStream_t stream)
{
	if (n == 0) return 0;
	LinesArgs a;
	memset(&a, 0, sizeof a);
	a.blob = static_cast<const uint8_t *>(dfa->d_lblob);
	a.blob_bytes = dfa->lblob_bytes;
	a.pitch = dfa->l_pitch;
	a.end_off = dfa->l_end_off;
	a.dead = dfa->l_dead;
	a.start = dfa->l_start;
	a.perm_inv = dfa->d_lperm_inv;
	a.absorb = getenv("FSM_B200_NO_ABSORB_SKIP") == nullptr ? dfa->d_labsorb : nullptr;
	a.masks = dfa->d_eager_masks;
	a.base = d_base; a.offsets = d_offsets; a.stride = stride; a.len = len; a.n = n;
	a.out = d_out; a.out_masks = d_masks;
	const uint32_t words = d_masks != nullptr ? dfa->eager_words : 0u;
	/* without a mask buffer only a missing edge is an event: the dead row is the last one */
	a.first_event = words != 0 ? dfa->l_first_event : dfa->l_dead;
	for (uint32_t w = 0; w < words && w < 4; w++) a.start_mask[w] = dfa->l_start_mask[w];
	a.ev_masks = dfa->d_lev_masks;
	a.prefetch = 8192;
	if (const char *e = getenv("FSM_B200_LINES_PREFETCH")) {        /* tuning knob: L2 prefetch distance, 0 = off */
		const int v = atoi(e);
		if (v >= 0 && v <= (1 << 20) && (v % 128) == 0) a.prefetch = (uint32_t) v;
	}
	if (dfa->l_entry_bytes == 1) return dispatch_lines<uint8_t>(a, words, !dfa->complete, dfa->device, stream);
	return dispatch_lines<uint16_t>(a, words, !dfa->complete, dfa->device, stream);
}

} // namespace fsmb200
