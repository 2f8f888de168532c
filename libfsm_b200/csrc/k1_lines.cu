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
 * The walk costs PRMT + LDS.U8 + LEA + LDS.U16 per byte, with no base-address arithmetic:
 *   - the table rows are indexed by byte class (16-bit entries) and after staging every entry is
 *     rewritten in shared memory from "next state" to that state's HANDLE = shared-memory address
 *     of its row / 4, so the dependent step is  st = lds16(4 * st + 2 * class)  -- one LEA;
 *   - the class LUT sits at a 512-aligned shared-memory address below 32 KiB: its address bits are
 *     OR-ed into the per-word validity bytes, and the PRMT that extracts byte t of a word builds the
 *     complete LUT address {byte, address bits | outside-the-line flag, 0, 0};
 *   - rows have one extra NOP column in which every state loops to itself.  A byte outside the
 *     line has its value cleared and the flag set: LUT[256] = NOP (one word for all such lanes: no
 *     bank conflict with the lanes inside their lines).  Words entirely outside the line are
 *     skipped;
 *   - dfa_compile.cu numbers the states with eager outputs last, just before the dead row, so
 *     "has ids" is one compare on the handle (handles are monotonic in the state number):
 *       no eager outputs (EV_DEAD): per byte the kernel adds max(st, dead - 1): the dead row absorbs,
 *         so after the sector the sum counts the steps spent dead and gives the offset of the missing
 *         edge; the state it was taken from is 0..3 exact steps away from a per-word snapshot;
 *       eager outputs (EV_EAGER): the four states entered in a word are kept; if their maximum has ids
 *         (1.75 % of the words of BASELINE config 3) the ids of that state -- and, rarely, of the other
 *         three -- are OR-ed in from a small global array.  OR is idempotent, so the NOP steps need no
 *         care.  (The first version accounted per sector and re-walked sectors with more than one
 *         event out of line: 0.5 % of the sectors, but 10 % of the warp iterations had such a lane, on
 *         its own for 32 serial steps -- 17 % of all executed instructions.)  Only a walk that ends in
 *         the dead row still takes the out-of-line exact walk, for the offset.
 *
 * Algorithmic bytes per line: its bytes, read once, + 16 B record (+ 8 W B id bitset).  Bound:
 * the shared-memory lookup rate (one conflict-free + one dependent, bank-conflicting LDS per byte),
 * see DESIGN.md.
 */
#include <cstring>

#include "common.h"
#include "k1_exec_batch.h"
#include "k1_device.cuh"

using namespace fsmb200;

namespace {

struct LinesArgs {
	const uint8_t *blob;          /* [516 B LUT, padded to 1024][rows][is_end] */
	uint32_t blob_bytes, pitch /* bytes, multiple of 4 */, tab_off, end_off, ntable, ncols;
	uint32_t first_event, dead, start;   /* state numbers (new numbering); NO_EDGE when absent */
	const uint32_t *perm_inv;     /* new state number -> caller's */
	const uint8_t *absorb;        /* by new number, or nullptr */
	const uint64_t *masks;        /* [ntable][W] by the caller's numbering (global memory) */
	const uint64_t *ev_masks;     /* [ntable - first_event][W] by new number - first_event */
	uint64_t start_mask[4];
	const uint8_t *base;
	const uint64_t *offsets;      /* n + 1 entries, or nullptr: fixed stride */
	const uint32_t *entry;        /* per line: the state its walk starts in (caller's numbering), or nullptr: the start state */
	const uint32_t *perm;         /* caller's state number -> new number (with entry) */
	uint64_t stride, len, n;
	fsm_b200_result *out;
	uint64_t *out_masks;          /* [n][W] */
	uint32_t l2_hint;             /* tuning: sector loads ask L2 for 256 B around a miss (neighbouring lanes' lines) */
};

/* the staged blob; file scope so that the out-of-line re-walk addresses it as shared memory too */
extern __shared__ __align__(1024) uint8_t lines_smem[];

/* {x.byte T, v.byte T, 0, 0}: PRMT with the sign-fill mode for the two upper bytes (v.byte T < 0x80) */
template <int T>
__device__ __forceinline__ uint32_t
byte_and_flag(uint32_t x, uint32_t v)
{
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(v), "n"(T | ((4 + T) << 4) | ((0xC + T) << 8) | ((0xC + T) << 12)));
	return d;
}

__device__ __forceinline__ uint32_t
lds_u8(uint32_t addr)
{
	uint32_t v;
	asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}

__device__ __forceinline__ uint32_t
lds_u16(uint32_t addr)
{
	uint32_t v;
	asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}

/* handle (row address / 4) <-> state number */
struct Handles {
	uint32_t tab4;       /* shared address of row 0, / 4 */
	uint32_t pitch4;     /* row pitch / 4 */
	uint32_t magic;      /* ceil(2^32 / pitch4): (x * magic) >> 32 == x / pitch4 for x < 2^16 */
	__device__ __forceinline__ uint32_t of(uint32_t state) const { return tab4 + state * pitch4; }
	__device__ __forceinline__ uint32_t state(uint32_t h) const { return __umulhi(h - tab4, magic); }
};

template <int W> struct Rewalk {
	uint32_t st, at, died;        /* st: handle */
	uint64_t m[W > 0 ? W : 1];
};

/* Exact walk of one sector: first byte (within `mask`) without an edge and the state it was taken
 * from, ids of every state entered.  Cold path; st / dead / first_event are handles. */
template <int W>
__device__ __noinline__ Rewalk<W>
lines_rewalk(Handles hd, uint32_t lut_a, uint32_t entry, uint32_t mask,
	uint32_t first_event, uint32_t dead, const uint32_t *perm_inv, const uint64_t *masks, const uint32_t (&w)[8])
{
	Rewalk<W> r;
	r.st = entry; r.at = 32; r.died = 0;
#pragma unroll
	for (int k = 0; k < (W > 0 ? W : 1); k++) r.m[k] = 0;
#pragma unroll 1
	for (uint32_t j = 0; j < 32; j++) {
		if (!((mask >> j) & 1u)) continue;
		const uint32_t b = (w[j >> 2] >> (8u * (j & 3u))) & 0xFFu;
		const uint32_t nx = lds_u16(4u * r.st + lds_u8(lut_a + b));
		if (nx == dead) { r.died = 1; r.at = j; break; }
		r.st = nx;
		if (W > 0 && nx >= first_event) {
			const uint32_t old = __ldg(perm_inv + hd.state(nx));
#pragma unroll
			for (int k = 0; k < W; k++) r.m[k] |= __ldg(masks + (size_t) old * W + k);
		}
	}
	return r;
}

constexpr int LINES_THREADS = 768;

enum { EV_NONE = 0, EV_DEAD = 1, EV_EAGER = 2 };

template <int W, int EV>
__global__ void __launch_bounds__(LINES_THREADS, 1)
k1_lines_kernel(const LinesArgs a)
{
	__shared__ uint64_t blob_bar;
	stage_blob(lines_smem, a.blob, a.blob_bytes, &blob_bar);

	const uint32_t lut_a = smem_u32(lines_smem);          /* 1024-aligned; must be < 32 KiB for the PRMT trick */
	Handles hd;
	hd.tab4 = (lut_a + a.tab_off) >> 2;
	hd.pitch4 = a.pitch >> 2;
	hd.magic = 0xFFFFFFFFu / hd.pitch4 + 1u;
	if (lut_a >= 0x8000u) __trap();                        /* static shared memory grew past 31 KiB: cannot happen in this file */
	{
		/* state numbers -> handles, in place (every CTA, once) */
		uint16_t *ent = reinterpret_cast<uint16_t *>(lines_smem + a.tab_off);
		const uint32_t nent = a.ntable * (a.pitch >> 1);
		for (uint32_t e = threadIdx.x; e < nent; e += blockDim.x) ent[e] = (uint16_t) hd.of(ent[e]);
		__syncthreads();
	}
	const uint32_t is_end_a = lut_a + a.end_off;
	const uint32_t lut_bits = (lut_a >> 8) * 0x01010101u;  /* address byte 1 of the LUT, in every byte */
	const uint32_t h_first = (EV != EV_NONE) ? hd.of(a.first_event) : 0xFFFFFFFFu;
	const uint32_t h_dead = a.dead != NO_EDGE ? hd.of(a.dead) : 0xFFFFFFFFu;
	const uint32_t h_start = hd.of(a.start);

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
			if (a.l2_hint) ld256_l2_256(reinterpret_cast<const uint8_t *>(saddr), w);
			else ld256(reinterpret_cast<const uint8_t *>(saddr), w);
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
	auto first_state = [&](uint64_t line) -> uint32_t {
		return a.entry != nullptr ? hd.of(__ldg(a.perm + __ldg(a.entry + line))) : h_start;
	};
	uint32_t st = have ? first_state(i) : h_start;
	uintptr_t line_beg = cur;
	uint64_t acc[W > 0 ? W : 1];
#pragma unroll
	for (int k = 0; k < (W > 0 ? W : 1); k++) acc[k] = W > 0 ? a.start_mask[k] : 0;

	/* every lane stays in the loop until the whole warp is done: the vote reconverges the warp at the
	 * top of every iteration (lanes finish and switch lines at different iterations; without it the
	 * groups that diverged there ran the walk separately from then on -- ncu: every walk instruction
	 * executed twice per iteration with two thirds of the lanes) */
	while (__any_sync(0xFFFFFFFFu, have)) {
		if (!have) continue;
		const uintptr_t saddr = cur & ~(uintptr_t) 31;
		const uint32_t lo = (uint32_t) (cur - saddr);
		const uint32_t hi = (end - saddr) >= 32 ? 32u : (uint32_t) (end - saddr);
		const bool more = saddr + 32 < end;              /* the line continues in the next sector */
		if (more) load_sector(saddr + 32, B);

		/* walk the sector; bytes outside [lo, hi) take the NOP column, words entirely outside are skipped */
		const uint32_t mask = (hi > lo) ? ((0xFFFFFFFFu << lo) & (0xFFFFFFFFu >> (32u - hi))) : 0u;
		const uint32_t inv = ~mask;
		const uint32_t entry = st;
		const uint32_t base_ev = h_first - 1u;            /* EV_DEAD: first_event is the dead row */
		uint32_t ssum = 0, nsteps = 0;
		uint32_t snap[8];
#pragma unroll
		for (int k = 0; k < 8; k++) {
			const uint32_t nib = (inv >> (4 * k)) & 0xFu;
			if (EV == EV_DEAD) snap[k] = st;
			if (nib != 0xFu) {
				/* byte t of out01 = 1 when byte 4k + t is outside the line; such bytes read LUT[256] */
				const uint32_t out01 = (nib * 0x00204081u) & 0x01010101u;
				const uint32_t x = A[k] & ~(out01 * 0xFFu);
				const uint32_t vw = out01 | lut_bits;
#define LINES_STEP(T)                                                                             \
				st = lds_u16(4u * st + lds_u8(byte_and_flag<T>(x, vw)));                          \
				if (EV == EV_DEAD) ssum += max(st, base_ev);
				if (EV == EV_EAGER) {
					/* the four states entered in this word; ids fire per word (OR is idempotent, so the
					 * NOP steps of bytes outside the line -- which re-enter the same state -- are harmless) */
					const uint32_t s0 = lds_u16(4u * st + lds_u8(byte_and_flag<0>(x, vw)));
					const uint32_t s1 = lds_u16(4u * s0 + lds_u8(byte_and_flag<1>(x, vw)));
					const uint32_t s2 = lds_u16(4u * s1 + lds_u8(byte_and_flag<2>(x, vw)));
					const uint32_t s3 = lds_u16(4u * s2 + lds_u8(byte_and_flag<3>(x, vw)));
					st = s3;
					const uint32_t m = max(max(s0, s1), max(s2, s3));
					if (m >= h_first) {
						/* handles grow with the state number: [h_first, h_dead) are the states with ids */
						if (m < h_dead) {
							const uint32_t ev = hd.state(m) - a.first_event;
#pragma unroll
							for (int q = 0; q < W; q++) acc[q] |= __ldg(a.ev_masks + (size_t) ev * W + q);
						}
						const bool others = (s0 >= h_first && s0 != m) || (s1 >= h_first && s1 != m) ||
						    (s2 >= h_first && s2 != m) || (s3 >= h_first && s3 != m);
						if (others) {
							const uint32_t sj[4] = { s0, s1, s2, s3 };
#pragma unroll
							for (int j = 0; j < 4; j++) {
								if (sj[j] >= h_first && sj[j] < h_dead && sj[j] != m) {
									const uint32_t ev = hd.state(sj[j]) - a.first_event;
#pragma unroll
									for (int q = 0; q < W; q++) acc[q] |= __ldg(a.ev_masks + (size_t) ev * W + q);
								}
							}
						}
					}
				} else {
					LINES_STEP(0) LINES_STEP(1) LINES_STEP(2) LINES_STEP(3)
				}
#undef LINES_STEP
				if (EV == EV_DEAD) nsteps += 4;
				if (EV == EV_DEAD) snap[k] = st;
			}
		}
		bool died = false;
		uint32_t consumed_here = hi - lo;
		if (EV == EV_DEAD && st == h_dead) {
			/* the dead row absorbs (NOP steps included): ssum - nsteps * base = steps spent dead.  Steps
			 * are taken in walked words only; map the count back to a sector byte index. */
			uint32_t dead_steps = ssum - nsteps * base_ev;
			uint32_t j = 32;                                  /* sector byte that had no edge */
#pragma unroll
			for (int k = 7; k >= 0; k--) {
				const uint32_t nib = (inv >> (4 * k)) & 0xFu;
				if (nib != 0xFu && dead_steps != 0) {
					if (dead_steps <= 4u) { j = 4u * (uint32_t) k + (4u - dead_steps); dead_steps = 0; }
					else dead_steps -= 4u;
				}
			}
			const uint32_t kq = j >> 2;
			uint32_t from = entry, word = A[0];
#pragma unroll
			for (int k = 1; k < 8; k++) { if (kq == (uint32_t) k) { from = snap[k - 1]; word = A[k]; } }
#pragma unroll
			for (int t = 0; t < 3; t++) {
				if ((uint32_t) t < (j & 3u) && ((mask >> (4u * kq + (uint32_t) t)) & 1u)) {
					from = lds_u16(4u * from + lds_u8(lut_a + ((word >> (8 * t)) & 0xFFu)));
				}
			}
			st = from;
			died = true;
			consumed_here = j - lo;
		}
		if (EV == EV_EAGER && st == h_dead) {
			/* a byte of this sector had no edge (the dead row absorbs): the exact walk finds which one and
			 * the state it was taken from; the ids it collects were OR-ed in above already */
			const Rewalk<W> r = lines_rewalk<W>(hd, lut_a, entry, mask, h_first, h_dead, a.perm_inv, a.masks, A);
			st = r.st;
#pragma unroll
			for (int k = 0; k < W; k++) acc[k] |= r.m[k];
			if (r.died) { died = true; consumed_here = r.at - lo; }
		}
		/* A state whose 256 edges all loop back to itself keeps the walk where it is whatever
		 * follows: the rest of the line cannot change the record (nor fire a new id), so it is
		 * neither walked nor read. */
		bool absorbed = false;
		if (!died && more && a.absorb != nullptr && __ldg(a.absorb + hd.state(st))) {
			absorbed = true;
			cur = end;
		} else {
			cur += consumed_here;
		}
		if (died || !more || absorbed) {
			/* line done */
			const uint32_t sn = hd.state(st);
			uint4 v;
			v.x = (!died && lds_u8(is_end_a + sn)) ? 1u : 0u;
			v.y = __ldg(a.perm_inv + sn);
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
				st = first_state(i);
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

template <int W, int EV>
int
launch_lines(const LinesArgs &a, int device, cudaStream_t stream)
{
	auto kern = k1_lines_kernel<W, EV>;
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

int
dispatch_lines(const LinesArgs &a, uint32_t words, bool dead, int device, cudaStream_t stream)
{
	switch (words) {
	case 0: return dead ? launch_lines<0, EV_DEAD>(a, device, stream) : launch_lines<0, EV_NONE>(a, device, stream);
	case 1: return launch_lines<1, EV_EAGER>(a, device, stream);
	case 2: return launch_lines<2, EV_EAGER>(a, device, stream);
	case 3: return launch_lines<3, EV_EAGER>(a, device, stream);
	case 4: return launch_lines<4, EV_EAGER>(a, device, stream);
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
	uint64_t stride, uint64_t len, size_t n, fsm_b200_result *d_out, uint64_t *d_masks, cudaStream_t stream,
	const uint32_t *d_entry)
{
	if (n == 0) return 0;
	LinesArgs a;
	memset(&a, 0, sizeof a);
	a.blob = static_cast<const uint8_t *>(dfa->d_lblob);
	a.blob_bytes = dfa->lblob_bytes;
	a.pitch = dfa->l_pitch;
	a.tab_off = dfa->l_tab_off;
	a.end_off = dfa->l_end_off;
	a.ntable = dfa->ntable;
	a.ncols = dfa->l_ncols;
	a.dead = dfa->l_dead;
	a.start = dfa->l_start;
	a.perm_inv = dfa->d_lperm_inv;
	a.absorb = getenv("FSM_B200_NO_ABSORB_SKIP") == nullptr ? dfa->d_labsorb : nullptr;
	if (const char *e = getenv("FSM_B200_LINES_L2HINT")) a.l2_hint = atoi(e) != 0 ? 1u : 0u;   /* tuning knob */
	a.masks = dfa->d_eager_masks;
	a.ev_masks = dfa->d_lev_masks;
	a.base = d_base; a.offsets = d_offsets; a.stride = stride; a.len = len; a.n = n;
	a.entry = d_entry; a.perm = dfa->d_lperm;
	a.out = d_out; a.out_masks = d_masks;
	const uint32_t words = d_masks != nullptr ? dfa->eager_words : 0u;
	/* without a mask buffer only a missing edge is an event: the dead row is the last one */
	a.first_event = words != 0 ? dfa->l_first_event : dfa->l_dead;
	for (uint32_t w = 0; w < words && w < 4; w++) a.start_mask[w] = dfa->l_start_mask[w];
	return dispatch_lines(a, words, !dfa->complete, dfa->device, stream);
}

} // namespace fsmb200
