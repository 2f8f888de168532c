/*
 * k1b_stream.cu -- K1b: ONE long input == one reference fsm_exec call over a stream
 * (src/libfsm/exec.c:132-151 is strictly serial: O(1) state carried byte to byte).
 *
 * DFA execution is a monoid: a byte range maps every entry state to an exit state, and maps
 * compose associatively.  The stream is cut into chunks of C bytes and processed in five
 * data-parallel steps, exact for ANY DFA (no reliance on self-synchronisation):
 *
 *   1. prefix   thread (chunk c, state s): walk the first W bytes of c from s
 *               -> img[c][s] (or "died": offset + state it died in).  Chains from wrong entry
 *               states die or merge within a few bytes for practical DFAs, so few distinct
 *               live images survive per chunk (1 for the config-2 and UTF-8 DFAs).  The first
 *               thread to reach a live image (c, v) claims it (atomicCAS) and appends ONE job:
 *               the rest of chunk c walked from v.
 *   2. body     the K1 kernels over the jobs (one lane per job, 256-bit loads, table in shared
 *               memory; k-stride when the DFA has one) -- this is where the bytes are scanned,
 *               at K1 speed.  The job count stays on the device (no host round trip).
 *   3. compose  two-level composition of the chunk maps in chunk order.  Level 1 (one block
 *               per group of G chunks) builds next[c][s] = exit of job(c, img[c][s]) on the fly
 *               in shared memory and folds the group, one thread per entry state; level 2
 *               folds the groups.  First-dead offset / dead-from state are resolved lazily.
 *
 * The result is the map entry-state -> (exit state | first dead offset + dead-from state)
 * of the whole range: exec_stream uses entry = start; the multi-GPU shard form all-gathers
 * the maps of the ranks' byte ranges and composes them in rank order (sharding.py).
 *
 * Applies to every table form up to 65534 states: 8-bit dense rows in shared memory use the prefix
 * kernel below, everything else (16-/32-bit entries, class-indexed rows, L2-resident tables) the
 * generic prefix kernel that reads the table through the read-only path; the body is whatever K1 kernel
 * the table form gets for a batch.  Inputs too short to cut run as a K1 batch of one (one lane).
 */
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "common.h"
#include "k1_exec_batch.h"
#include "k1b_rep.cuh"
#include "k1b_rep_tma.cuh"

using namespace fsmb200;

namespace {

constexpr uint32_t DEADMARK = 0xFFFFFFFFu;
constexpr uint16_t DEAD16 = 0xFFFFu;
constexpr uint32_t PREFIX_W_MAX = 64;
constexpr uint32_t NOJOB = 0xFFFFFFFFu;
constexpr uint32_t ABSORBED = 0xFFFFFFFDu;     /* job_of marker: image state is absorbing, no scan needed */

struct StreamArgs {
	const uint8_t *buf;
	uint64_t len;
	uint64_t C;            /* chunk bytes */
	uint32_t nchunks;
	uint32_t T;            /* real states (entry states considered) */
	uint32_t W;            /* prefix window (bytes) */
	const uint8_t *absorb; /* [ntable] */
	const uint8_t *blob;
	uint32_t blob_bytes, pitch, dead;
	uint32_t entry_bytes, cls_off, has_cls;   /* generic prefix kernel: table form (global memory) */
	/* per (chunk, state) */
	uint32_t *img, *pdo, *pdf;
	uint32_t *job_of;      /* [chunk][image state] -> job index; NOJOB / PENDING while unclaimed */
	/* jobs */
	uint64_t *job_beg, *job_end;
	uint32_t *job_entry;
	uint32_t *njobs;
	const fsm_b200_result *rec;
};

__device__ __forceinline__ uint32_t
smem_u32(const void *p)
{
	return (uint32_t) __cvta_generic_to_shared(p);
}

/* Stage the table blob with TMA bulk copies (same scheme as K1). */
__device__ __forceinline__ void
stage_blob(uint8_t *smem, const uint8_t *blob, uint32_t blob_bytes, uint64_t *bar)
{
	const uint32_t bar_a = smem_u32(bar);
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(blob_bytes) : "memory");
		const uint32_t dst = smem_u32(smem);
		for (uint32_t off = 0; off < blob_bytes; off += 16384u) {
			const uint32_t nb = min(16384u, blob_bytes - off);
			asm volatile(
			    "cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			    :: "r"(dst + off), "l"(blob + off), "r"(nb), "r"(bar_a) : "memory");
		}
	}
	uint32_t done;
	do {
		asm volatile(
		    "{\n\t.reg .pred p;\n\t"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		    "selp.u32 %0, 1, 0, p;\n\t}"
		    : "=r"(done) : "r"(bar_a), "r"(0) : "memory");
	} while (!done);
}

/* 1. prefix: thread per (chunk, entry state).  Lanes of a warp share the chunk (same bytes:
 * broadcast loads, independent of the state so the unrolled loop batches them) and hold
 * consecutive states (row pitch 260 B: distinct banks).  The dead row absorbs, so there is
 * no early exit; the first death is recorded with selects. */
__global__ void __launch_bounds__(1024, 2)
k1b_prefix_kernel(const StreamArgs a)
{
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint64_t blob_bar;
	stage_blob(smem, a.blob, a.blob_bytes, &blob_bar);

	const uint64_t total = (uint64_t) a.nchunks * a.T;
	const uint64_t nthreads = (uint64_t) gridDim.x * blockDim.x;
	for (uint64_t idx = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += nthreads) {
		const uint32_t c = (uint32_t) (idx / a.T), s = (uint32_t) (idx % a.T);
		const uint64_t beg = (uint64_t) c * a.C;
		const uint64_t clen = min(a.C, a.len - beg);
		const uint32_t w = (uint32_t) min((uint64_t) a.W, clen);
		const uint8_t *p = a.buf + beg;
		uint32_t st = s, dk = 0, df = s;
		bool died = false;
		if (w == a.W) {
#pragma unroll 16
			for (uint32_t k = 0; k < a.W; k++) {
				const uint32_t nx = smem[st * a.pitch + __ldg(p + k)];
				const bool hit = !died && nx == a.dead;
				dk = hit ? k : dk;
				df = hit ? st : df;
				died = died || hit;
				st = nx;
			}
		} else {
			for (uint32_t k = 0; k < w; k++) {
				const uint32_t nx = smem[st * a.pitch + __ldg(p + k)];
				if (nx == a.dead) { died = true; dk = k; df = st; break; }
				st = nx;
			}
		}
		if (died) {
			a.img[idx] = DEADMARK; a.pdo[idx] = dk; a.pdf[idx] = df;
		} else {
			a.img[idx] = st;
			/* first thread to produce the live image (c, st) appends its job */
			uint32_t *slot = &a.job_of[(uint64_t) c * a.T + st];
			if (a.absorb[st]) {
				*slot = ABSORBED;               /* leaves the chunk as it entered: no job */
			} else if (atomicCAS(slot, NOJOB, NOJOB - 1u) == NOJOB) {
				const uint32_t j = atomicAdd(a.njobs, 1u);
				a.job_beg[j] = beg + w;
				a.job_end[j] = beg + clen;
				a.job_entry[j] = st;
				*slot = j;                      /* read by the compose kernels (later launches) */
			}
		}
	}
}

/* 1'. prefix for every other table form (16- / 32-bit entries, class-indexed rows, tables too large
 * for shared memory): the same walk with the table read from global memory through the read-only
 * path (the rows touched by W steps from consecutive entry states are L1/L2-resident).  T x W lookups
 * per chunk -- next to a chunk of at least 128 x T bytes that is under 1/8 lookup per input byte. */
template <typename E>
__global__ void __launch_bounds__(256)
k1b_prefix_generic_kernel(const StreamArgs a)
{
	const uint64_t total = (uint64_t) a.nchunks * a.T;
	const uint64_t nthreads = (uint64_t) gridDim.x * blockDim.x;
	const uint8_t *cls = a.blob + a.cls_off;
	for (uint64_t idx = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += nthreads) {
		const uint32_t c = (uint32_t) (idx / a.T), s = (uint32_t) (idx % a.T);
		const uint64_t beg = (uint64_t) c * a.C;
		const uint64_t clen = min(a.C, a.len - beg);
		const uint32_t w = (uint32_t) min((uint64_t) a.W, clen);
		const uint8_t *p = a.buf + beg;
		uint32_t st = s, dk = 0, df = s;
		bool died = false;
		for (uint32_t k = 0; k < w; k++) {
			const uint32_t b = __ldg(p + k);
			const uint32_t col = a.has_cls ? (uint32_t) __ldg(cls + b) : b;
			const uint32_t nx = (uint32_t) __ldg(reinterpret_cast<const E *>(a.blob + (size_t) st * a.pitch) + col);
			if (nx == a.dead) { died = true; dk = k; df = st; break; }
			st = nx;
		}
		if (died) {
			a.img[idx] = DEADMARK; a.pdo[idx] = dk; a.pdf[idx] = df;
		} else {
			a.img[idx] = st;
			uint32_t *slot = &a.job_of[(uint64_t) c * a.T + st];
			if (a.absorb[st]) {
				*slot = ABSORBED;
			} else if (atomicCAS(slot, NOJOB, NOJOB - 1u) == NOJOB) {
				const uint32_t j = atomicAdd(a.njobs, 1u);
				a.job_beg[j] = beg + w;
				a.job_end[j] = beg + clen;
				a.job_entry[j] = st;
				*slot = j;
			}
		}
	}
}

/* The chunk map, evaluated lazily: where does chunk c take entry state s?
 * Returns the exit state, or DEAD16 with the global first-dead offset and dead-from state. */
__device__ __forceinline__ uint16_t
chunk_next(const StreamArgs &a, uint32_t c, uint32_t s, uint64_t *dead_off, uint32_t *dead_from)
{
	const uint64_t idx = (uint64_t) c * a.T + s;
	const uint32_t v = a.img[idx];
	if (v == DEADMARK) {
		if (dead_off) { *dead_off = (uint64_t) c * a.C + a.pdo[idx]; *dead_from = a.pdf[idx]; }
		return DEAD16;
	}
	const uint32_t j = a.job_of[(uint64_t) c * a.T + v];
	if (j == ABSORBED) return (uint16_t) v;
	const fsm_b200_result r = a.rec[j];
	const uint64_t jb = a.job_beg[j], jlen = a.job_end[j] - jb;
	if (r.consumed < jlen) {                 /* the body walk hit a missing edge */
		if (dead_off) { *dead_off = jb + r.consumed; *dead_from = r.end; }
		return DEAD16;
	}
	return (uint16_t) r.end;
}

/* 3. compose, level 1: block b folds chunks [b*G, b*G+G) for every entry state.
 * gnext[b][s] = exit or DEAD16; on death gchunk/gstate say where (chunk, state entering it). */
__global__ void
k1b_compose1_kernel(const StreamArgs a, uint32_t G, uint16_t *gnext, uint32_t *gchunk, uint32_t *gstate)
{
	extern __shared__ __align__(16) uint16_t sm[];
	const uint32_t T = a.T;
	const uint32_t c0 = blockIdx.x * G;
	const uint32_t cn = min(G, a.nchunks - c0);
	for (uint32_t i = threadIdx.x; i < cn * T; i += blockDim.x) {
		sm[i] = chunk_next(a, c0 + i / T, i % T, nullptr, nullptr);
	}
	__syncthreads();
	for (uint32_t s = threadIdx.x; s < T; s += blockDim.x) {
		uint32_t st = s, dc = 0xFFFFFFFFu, ds = 0;
		for (uint32_t k = 0; k < cn; k++) {
			const uint16_t nx = sm[k * T + st];
			if (nx == DEAD16) { dc = c0 + k; ds = st; st = DEAD16; break; }
			st = nx;
		}
		gnext[(uint64_t) blockIdx.x * T + s] = (uint16_t) st;
		gchunk[(uint64_t) blockIdx.x * T + s] = dc;
		gstate[(uint64_t) blockIdx.x * T + s] = ds;
	}
}

/* 3. compose, level 2: one thread per entry state folds the group maps in order. */
__global__ void
k1b_compose2_kernel(const StreamArgs a, const uint16_t *gnext, const uint32_t *gchunk, const uint32_t *gstate,
	uint32_t ngroups, uint32_t smem_groups, StreamOut *out)
{
	extern __shared__ __align__(16) uint16_t sm[];
	const uint32_t T = a.T;
	const bool staged = smem_groups != 0;            /* single block: group maps fit shared memory */
	if (staged) {
		for (uint64_t i = threadIdx.x; i < (uint64_t) ngroups * T; i += blockDim.x) sm[i] = gnext[i];
		__syncthreads();
	}
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= T) return;
	uint32_t st = s;
	for (uint32_t g = 0; g < ngroups; g++) {
		const uint64_t i = (uint64_t) g * T + st;
		const uint16_t nx = staged ? sm[i] : gnext[i];
		if (nx == DEAD16) {
			uint64_t off = 0; uint32_t from = 0;
			(void) chunk_next(a, gchunk[i], gstate[i], &off, &from);
			out[s].state = from;
			out[s].died = 1;
			out[s].dead_off = off;
			return;
		}
		st = nx;
	}
	out[s].state = st;
	out[s].died = 0;
	out[s].dead_off = 0xFFFFFFFFFFFFFFFFull;
}

/* ------------------------------------------------------------------ host side -------- */

struct Arena {
	uint8_t *base = nullptr;
	size_t cap = 0, used = 0;
	template <typename T> T *take(size_t n) {
		used = (used + 255) & ~(size_t) 255;
		T *p = reinterpret_cast<T *>(base + used);
		used += n * sizeof(T);
		return p;
	}
};

struct StreamScratch {
	std::mutex mu;
	Arena arena;
	uint8_t *d_in = nullptr; size_t in_cap = 0;       /* for the _host entry point */
	uint8_t *d_eager = nullptr; size_t eager_cap = 0; /* eager streams: per-chunk offsets, entry states, records, id masks */
	cudaStream_t stream = nullptr;
	/* small automata (k1b_rep.cuh): dense [ntable][256] next-state bytes, absorbing states, output arena */
	uint8_t *d_dense = nullptr;
	uint32_t absorb_mask = 0;
	Arena rep_arena;
	StreamOut *h_pinned = nullptr;                    /* [REP_MAX_ROWS] read-back buffer */
	int sms = 0;                                      /* per-call driver queries are a measurable share of a 1 MiB call */
};

/* Opt a kernel in to the device's maximum dynamic shared memory ONCE per (device, kernel): the attribute is a
 * per-function maximum, so it is set to the device limit (never to one automaton's need, which a smaller
 * automaton would lower again) and remembered. */
int
ensure_max_smem(const void *kern, int device)
{
	static std::mutex mu;
	static std::vector<std::pair<int, const void *>> done;
	std::lock_guard<std::mutex> g(mu);
	for (const auto &d : done) if (d.first == device && d.second == kern) return 0;
	int optin = 0;
	FSMB_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device), return -1);
	FSMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin), return -1);
	done.emplace_back(device, kern);
	return 0;
}

std::mutex g_ss_mu;

StreamScratch *
ss_get(const fsm_b200_dfa *cdfa)
{
	fsm_b200_dfa *dfa = const_cast<fsm_b200_dfa *>(cdfa);
	std::lock_guard<std::mutex> g(g_ss_mu);
	if (dfa->stream_scratch == nullptr) dfa->stream_scratch = new (std::nothrow) StreamScratch();
	return static_cast<StreamScratch *>(dfa->stream_scratch);
}

constexpr uint32_t K1B_MAX_STATES = 65534u;

/* 8-bit dense rows in shared memory: the round-1 prefix kernel */
bool
small_table(const fsm_b200_dfa *dfa)
{
	return dfa->smem_resident && dfa->nclasses == 0 && dfa->entry_bytes == 1;
}

uint32_t
pick_window(uint32_t T)
{
	/* chains from wrong entry states merge or die within a few bytes for practical DFAs; a
	 * shorter window for big tables keeps the T x W prefix walks cheap (a chain that has not
	 * merged yet only costs an extra job, never exactness) */
	uint32_t w = T <= 16 ? 64u : (T <= 64 ? 32u : 16u);
	if (const char *e = getenv("FSM_B200_STREAM_WINDOW")) {
		const int v = atoi(e);
		if (v >= 1 && v <= (int) PREFIX_W_MAX) w = (uint32_t) v;
	}
	return w;
}

size_t
pick_chunk(uint32_t T, uint32_t W, uint64_t len, int sms)
{
	/* One body job per lane and about two waves of lanes (measured on B200, 2 GiB of UTF-8:
	 * 2 KiB chunks 1.39 TB/s, 8 KiB 2.07 TB/s, 32 KiB 0.99 TB/s -- profiles/r1_k1b_stream.jsonl);
	 * never so small that the T x W prefix walks outweigh the body. */
	const uint64_t lanes = 2ull * (uint64_t) (sms > 0 ? sms : 148) * 1024ull;
	uint64_t lo = 2048;
	(void) W;
	while (lo < 128ull * T) lo <<= 1;      /* per-chunk bookkeeping (T map entries) must stay small next to the chunk */
	uint64_t c = lo;
	while (c < (1ull << 22) && len / c > lanes) c <<= 1;
	if (const char *e = getenv("FSM_B200_STREAM_CHUNK")) {
		const long v = atol(e);
		if (v >= 64 && v <= (1l << 30)) c = (uint64_t) v & ~31ull;
	}
	while (len / c > (1ull << 22)) c <<= 1;           /* bound the number of chunks */
	return (size_t) c;
}

/* ---- small automata: the fused form (k1b_rep.cuh) ------------------------------------------------ */

typedef CUresult (*rep_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

rep_encode_tiled_fn
rep_get_encode_tiled()
{
	static rep_encode_tiled_fn fn = nullptr;
	static std::once_flag once;
	std::call_once(once, [] {
		void *p = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
		    qres == cudaDriverEntryPointSuccess) {
			fn = reinterpret_cast<rep_encode_tiled_fn>(p);
		}
	});
	return fn;
}

/* The TMA-tile form (k1b_rep_tma.cuh): returns 1 when it launched, 0 when this call does not qualify (the
 * caller then launches the 256-bit-load form), -1 on error. */
int
launch_rep_tma(const fsm_b200_dfa *dfa, const RepArgs &a, unsigned grid, cudaStream_t stream)
{
	/* Opt-in: measured SLOWER than the 256-bit-load form (2 GiB of UTF-8: 1.24 ms with 4 stages, 0.96 with 3,
	 * 0.59 with 2, against 0.50 ms) -- tiles of 32-byte rows are the wrong shape for the TMA unit (at 4 stages one box row per
	 * 5.3 cycles per SM = 1.8 TB/s, 65 % of the warp time on the full barrier; profiles/r2_k1b_rep_tma_ncu_full.txt).  Kept selectable and parity-tested as the record of that. */
	const char *tma_env = getenv("FSM_B200_REP_TMA");
	if (tma_env == nullptr || atoi(tma_env) == 0) return 0;
	const uint64_t nfull64 = a.len / a.C;
	if (a.mis != 0 || nfull64 < 32 || nfull64 > 0x7FFFFFFFull || a.C > 0x7FFFFFFFull) return 0;
	int smem_optin = 0;
	if (cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dfa->device) != cudaSuccess) return 0;
	/* [maps + barriers][pad to an 8 KiB-aligned shared address (at most 8 KiB)][table][ring: 32 warps x NST KiB] */
	const size_t ring_off = ((size_t) REPC_HEAD_BYTES + 8192u + ((size_t) a.ntable << REPC_ROW_SHIFT) + 1023u) & ~(size_t) 1023u;
	int nst = (int) (((size_t) smem_optin - ring_off) / (32u * REPC_STAGE_BYTES));
	if ((size_t) smem_optin < ring_off) nst = 0;
	if (nst > 4) nst = 4;
	if (const char *e = getenv("FSM_B200_REP_TMA_STAGES")) { const int v = atoi(e); if (v >= 2 && v < nst) nst = v; }   /* tuning knob */
	if (nst < 2) return 0;
	rep_encode_tiled_fn enc = rep_get_encode_tiled();
	if (enc == nullptr) return 0;
	CUtensorMap tmap;
	const cuuint64_t gdim[2] = { (cuuint64_t) a.C, (cuuint64_t) nfull64 };
	const cuuint64_t gstr[1] = { (cuuint64_t) a.C };
	const cuuint32_t box[2] = { 32u, 32u };
	const cuuint32_t estr[2] = { 1u, 1u };
	const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(a.buf), gdim, gstr, box, estr,
	    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) return 0;
	const size_t smem = ring_off + (size_t) nst * 32u * REPC_STAGE_BYTES;
	const bool dead = !dfa->complete;
	void (*kern)(const RepArgs, const CUtensorMap, const uint32_t, const uint32_t);
	if (nst == 4) kern = dead ? k1b_rep_tma_kernel<true, 4> : k1b_rep_tma_kernel<false, 4>;
	else if (nst == 3) kern = dead ? k1b_rep_tma_kernel<true, 3> : k1b_rep_tma_kernel<false, 3>;
	else kern = dead ? k1b_rep_tma_kernel<true, 2> : k1b_rep_tma_kernel<false, 2>;
	FSMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), return -1);
	kern<<<grid, 1024, smem, stream>>>(a, tmap, (uint32_t) nfull64, (uint32_t) ring_off);
	return 1;
}

bool
rep_eligible(const fsm_b200_dfa *dfa)
{
	if (dfa->ntable > REP_MAX_ROWS || dfa->nstates == 0 || dfa->h_table32 == nullptr) return false;
	if (const char *e = getenv("FSM_B200_STREAM_REP")) return atoi(e) != 0;
	return true;
}

int
stream_map_rep(const fsm_b200_dfa *dfa, const uint8_t *d_buf, uint64_t len, cudaStream_t stream,
	StreamScratch *ss, std::vector<StreamOut> &h_out, StreamOut *d_user_out = nullptr)
{
	const uint32_t T = dfa->nstates, NT = dfa->ntable;
	if (ss->sms == 0) {
		FSMB_CUDA(cudaDeviceGetAttribute(&ss->sms, cudaDevAttrMultiProcessorCount, dfa->device), return -1);
	}
	const int sms = ss->sms;
	if (ss->d_dense == nullptr) {
		/* dense next-state bytes from the host copy of the table (missing edge -> dead row, which absorbs) */
		std::vector<uint8_t> dense((size_t) NT * 256);
		uint32_t mask = 0;
		for (uint32_t s = 0; s < NT; s++) {
			bool self = true;
			for (uint32_t b = 0; b < 256; b++) {
				uint32_t v = s < T ? dfa->h_table32[(size_t) s * 256 + b] : dfa->dead;
				if (v == NO_EDGE) v = dfa->dead;
				dense[(size_t) s * 256 + b] = (uint8_t) v;
				self = self && v == s;
			}
			if (self) mask |= 1u << s;
		}
		void *p = nullptr;
		FSMB_CUDA(cudaMalloc(&p, dense.size()), return -1);
		FSMB_CUDA(cudaMemcpy(p, dense.data(), dense.size(), cudaMemcpyHostToDevice), { cudaFree(p); return -1; });
		FSMB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&ss->h_pinned), REP_MAX_ROWS * sizeof(StreamOut)), { cudaFree(p); return -1; });
		ss->d_dense = static_cast<uint8_t *>(p);
		ss->absorb_mask = mask;
	}
	RepArgs a;
	memset(&a, 0, sizeof a);
	a.buf = d_buf; a.len = len;
	a.mis = (uint32_t) (reinterpret_cast<uintptr_t>(d_buf) & 31u);
	a.T = T; a.ntable = NT; a.dead = dfa->complete ? NO_EDGE : dfa->dead;
	a.absorb_mask = ss->absorb_mask; a.dense = ss->d_dense;
	/* one chunk per lane of ONE full wave; never so short that the T x 64-byte prefix walks dominate */
	const uint64_t lanes = (uint64_t) sms * 1024u;
	uint64_t C = ((len + a.mis + lanes - 1) / lanes + 31u) & ~31ull;
	uint64_t cmin = (uint64_t) T * REP_W;
	if (cmin < 256) cmin = 256;
	cmin = (cmin + 31u) & ~31ull;
	if (C < cmin) C = cmin;
	if (const char *e = getenv("FSM_B200_STREAM_CHUNK")) {
		const long v = atol(e);
		if (v >= 64 && v <= (1l << 30) && ((uint64_t) v & ~31ull) >= C) C = (uint64_t) v & ~31ull;   /* only larger: one wave */
	}
	a.C = C;
	a.nchunks = (uint32_t) ((len + a.mis + C - 1) / C);
	a.nwarps = (a.nchunks + 31u) / 32u;
	/* as many CTAs as there are SMs (or warps); CTA b walks the consecutive warps [b * wpc, b * wpc + wpc) */
	a.wpc = (a.nwarps + (uint32_t) sms - 1u) / (uint32_t) sms;
	a.nmaps = (a.nwarps + a.wpc - 1u) / a.wpc;
	const uint32_t nlv = (a.nmaps + 31u) / 32u;

	const size_t need = (size_t) a.nmaps * 16 * (1 + 4 + 1) + (size_t) a.nchunks * 16 * (8 + 1) +
	    (size_t) nlv * 16 * (4 + 1) * 2 + REP_MAX_ROWS * sizeof(StreamOut) + 16 * 256;
	if (ss->rep_arena.cap < need) {
		if (ss->rep_arena.base) cudaFree(ss->rep_arena.base);
		ss->rep_arena.base = nullptr; ss->rep_arena.cap = 0;
		void *p = nullptr;
		FSMB_CUDA(cudaMalloc(&p, need + need / 8), return -1);
		ss->rep_arena.base = static_cast<uint8_t *>(p);
		ss->rep_arena.cap = need + need / 8;
	}
	Arena &ar = ss->rep_arena;
	ar.used = 0;
	a.wmap = ar.take<uint8_t>((size_t) a.nmaps * 16);
	a.wdc = ar.take<uint32_t>((size_t) a.nmaps * 16);
	a.wds = ar.take<uint8_t>((size_t) a.nmaps * 16);
	a.body_off = ar.take<uint64_t>((size_t) a.nchunks * 16);
	a.body_from = ar.take<uint8_t>((size_t) a.nchunks * 16);
	for (int k = 0; k < 2; k++) {
		a.lv_dc[k] = ar.take<uint32_t>((size_t) nlv * 16);
		a.lv_ds[k] = ar.take<uint8_t>((size_t) nlv * 16);
	}
	a.out = ar.take<StreamOut>(REP_MAX_ROWS);
	if (d_user_out != nullptr) a.out = d_user_out;       /* asynchronous form: the records stay on the device */

	/* [chunk maps 8 KiB][pad to a 16 KiB-aligned shared address][table]; the pad is at most 16 KiB */
	const size_t smem = REP_MAPS_BYTES + 16384u + ((size_t) NT << REP_ROW_SHIFT);
	const unsigned grid = a.nmaps;
	/* FSM_B200_REP_TMA=1: input by TMA tiles (k1b_rep_tma.cuh; aligned buffers with at least one warp of full chunks) */
	const int tma = launch_rep_tma(dfa, a, grid, stream);
	if (tma < 0) return -1;
	if (tma == 0) {
		int hint = 1, nbuf = 4;      /* measured (profiles/r2_k1b_rep_knobs.jsonl): 2 buffers 0.60 ms, 3 0.56 ms, 4 0.50 ms per 2 GiB */
		if (const char *e = getenv("FSM_B200_REP_L2HINT")) hint = atoi(e) != 0;             /* tuning knobs */
		if (const char *e = getenv("FSM_B200_REP_NBUF")) { const int v = atoi(e); if (v >= 2 && v <= 4) nbuf = v; }
		void (*kern)(const RepArgs);
		if (nbuf == 4) {
			kern = dfa->complete ? (hint ? k1b_rep_kernel<false, 1, 4> : k1b_rep_kernel<false, 0, 4>)
			                     : (hint ? k1b_rep_kernel<true, 1, 4> : k1b_rep_kernel<true, 0, 4>);
		} else if (nbuf == 3) {
			kern = dfa->complete ? (hint ? k1b_rep_kernel<false, 1, 3> : k1b_rep_kernel<false, 0, 3>)
			                     : (hint ? k1b_rep_kernel<true, 1, 3> : k1b_rep_kernel<true, 0, 3>);
		} else {
			kern = dfa->complete ? (hint ? k1b_rep_kernel<false, 1, 2> : k1b_rep_kernel<false, 0, 2>)
			                     : (hint ? k1b_rep_kernel<true, 1, 2> : k1b_rep_kernel<true, 0, 2>);
		}
		if (ensure_max_smem(reinterpret_cast<const void *>(kern), dfa->device) != 0) return -1;
		kern<<<grid, 1024, smem, stream>>>(a);
	}
	count_launch();
	const size_t smem2 = (size_t) a.nmaps * 16 + (size_t) nlv * 16 + 16;
	if (smem2 > 48u * 1024u) {       /* more than 3000 CTA maps: not on this hardware, but the opt-in is needed then */
		FSMB_CUDA(cudaFuncSetAttribute(k1b_rep_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem2), return -1);
	}
	k1b_rep_final_kernel<<<1, 1024, smem2, stream>>>(a);
	count_launch();
	FSMB_CUDA(cudaGetLastError(), return -1);
	if (d_user_out != nullptr) return 0;
	FSMB_CUDA(cudaMemcpyAsync(ss->h_pinned, a.out, T * sizeof(StreamOut), cudaMemcpyDeviceToHost, stream), return -1);
	FSMB_CUDA(cudaStreamSynchronize(stream), return -1);
	h_out.assign(ss->h_pinned, ss->h_pinned + T);
	return 0;
}

/* Run steps 1-5 over d_buf[0..len); leaves StreamOut[T] in h_out. */
int
stream_map(const fsm_b200_dfa *dfa, const uint8_t *d_buf, uint64_t len, cudaStream_t stream,
	StreamScratch *ss, std::vector<StreamOut> &h_out, StreamArgs *keep = nullptr, StreamOut *d_user_out = nullptr)
{
	if (keep == nullptr && rep_eligible(dfa)) return stream_map_rep(dfa, d_buf, len, stream, ss, h_out, d_user_out);
	const uint32_t T = dfa->nstates;
	int sms = 0;
	FSMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dfa->device), return -1);
	const uint32_t W = pick_window(T);
	const uint64_t C = pick_chunk(T, W, len, sms);
	const uint32_t nchunks = (uint32_t) ((len + C - 1) / C);
	const uint64_t cs = (uint64_t) nchunks * T;
	/* level-1 groups: small enough that many blocks run (G <= 256) and G*T*2 B fits 32 KB */
	uint32_t G = 16384u / T;
	if (G > 256) G = 256;
	if (G < 1) G = 1;
	const uint32_t ngroups = (nchunks + G - 1) / G;

	const size_t need = cs * (4 + 4 + 4 + 4) + cs * (8 + 8 + 4 + 16) +
	    (size_t) ngroups * T * (2 + 4 + 4) + T * sizeof(StreamOut) + 64 * 256 + 8192;
	if (ss->arena.cap < need) {
		if (ss->arena.base) cudaFree(ss->arena.base);
		ss->arena.base = nullptr; ss->arena.cap = 0;
		void *p = nullptr;
		FSMB_CUDA(cudaMalloc(&p, need), return -1);
		ss->arena.base = static_cast<uint8_t *>(p);
		ss->arena.cap = need;
	}
	Arena &ar = ss->arena;
	ar.used = 0;

	StreamArgs a;
	memset(&a, 0, sizeof a);
	a.buf = d_buf; a.len = len; a.C = C; a.nchunks = nchunks; a.T = T; a.W = W; a.absorb = dfa->d_absorb;
	a.blob = static_cast<const uint8_t *>(dfa->d_blob);
	a.blob_bytes = (uint32_t) dfa->blob_bytes; a.pitch = dfa->pitch; a.dead = dfa->dead;
	a.img = ar.take<uint32_t>(cs); a.pdo = ar.take<uint32_t>(cs); a.pdf = ar.take<uint32_t>(cs);
	a.job_of = ar.take<uint32_t>(cs);
	a.job_beg = ar.take<uint64_t>(cs); a.job_end = ar.take<uint64_t>(cs); a.job_entry = ar.take<uint32_t>(cs);
	fsm_b200_result *rec = ar.take<fsm_b200_result>(cs);
	a.rec = rec;
	a.njobs = ar.take<uint32_t>(64);
	uint16_t *gnext = ar.take<uint16_t>((size_t) ngroups * T);
	uint32_t *gchunk = ar.take<uint32_t>((size_t) ngroups * T);
	uint32_t *gstate = ar.take<uint32_t>((size_t) ngroups * T);
	StreamOut *d_out = ar.take<StreamOut>(T);

	FSMB_CUDA(cudaMemsetAsync(a.job_of, 0xFF, cs * sizeof(uint32_t), stream), return -1);
	FSMB_CUDA(cudaMemsetAsync(a.njobs, 0, sizeof(uint32_t), stream), return -1);

	a.entry_bytes = dfa->entry_bytes; a.cls_off = dfa->cls_off; a.has_cls = dfa->nclasses != 0 ? 1u : 0u;
	if (small_table(dfa)) {
		const size_t smem_bytes = (dfa->blob_bytes + 127u) & ~(size_t) 127u;
		FSMB_CUDA(cudaFuncSetAttribute(k1b_prefix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_bytes), return -1);
		int per_sm = 1;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1b_prefix_kernel, 1024, smem_bytes) != cudaSuccess || per_sm < 1) per_sm = 1;
		uint64_t want = (cs + 1023) / 1024;
		const uint64_t cap = (uint64_t) sms * (uint64_t) per_sm;
		unsigned grid = (unsigned) (want < cap ? want : cap);
		k1b_prefix_kernel<<<grid, 1024, smem_bytes, stream>>>(a);
		count_launch();
	} else {
		uint64_t want = (cs + 255) / 256;
		const uint64_t cap = (uint64_t) sms * 8u;
		const unsigned grid = (unsigned) (want < cap ? want : cap);
		if (dfa->entry_bytes == 1) k1b_prefix_generic_kernel<uint8_t><<<grid, 256, 0, stream>>>(a);
		else if (dfa->entry_bytes == 2) k1b_prefix_generic_kernel<uint16_t><<<grid, 256, 0, stream>>>(a);
		else k1b_prefix_generic_kernel<uint32_t><<<grid, 256, 0, stream>>>(a);
		count_launch();
	}
	/* body: at most one job per (chunk, state); the actual count is read on the device */
	if (k1_launch_jobs(dfa, d_buf, a.job_beg, a.job_end, a.job_entry, cs, a.njobs, rec, stream) != 0) return -1;
	const size_t sm1 = (size_t) G * T * sizeof(uint16_t);
	FSMB_CUDA(cudaFuncSetAttribute(k1b_compose1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm1), return -1);
	k1b_compose1_kernel<<<ngroups, 256, sm1, stream>>>(a, G, gnext, gchunk, gstate);
	count_launch();
	{
		const size_t sm2 = (size_t) ngroups * T * sizeof(uint16_t);
		const bool staged = sm2 <= 160 * 1024 && T <= 1024;
		if (staged) {
			FSMB_CUDA(cudaFuncSetAttribute(k1b_compose2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm2), return -1);
			k1b_compose2_kernel<<<1, T < 32 ? 256 : ((T + 31) / 32) * 32 < 256 ? 256 : ((T + 31) / 32) * 32, sm2, stream>>>(a, gnext, gchunk, gstate, ngroups, 1u, d_out);
		} else {
			k1b_compose2_kernel<<<(T + 255) / 256, 256, 0, stream>>>(a, gnext, gchunk, gstate, ngroups, 0u, d_out);
		}
		count_launch();
	}
	FSMB_CUDA(cudaGetLastError(), return -1);
	if (keep != nullptr) *keep = a;      /* the chunk maps stay in the arena until the next call on this DFA */
	if (d_user_out != nullptr) {         /* asynchronous form: the records stay on the device, nobody waits */
		FSMB_CUDA(cudaMemcpyAsync(d_user_out, d_out, T * sizeof(StreamOut), cudaMemcpyDeviceToDevice, stream), return -1);
		return 0;
	}
	h_out.resize(T);
	FSMB_CUDA(cudaMemcpyAsync(h_out.data(), d_out, T * sizeof(StreamOut), cudaMemcpyDeviceToHost, stream), return -1);
	FSMB_CUDA(cudaStreamSynchronize(stream), return -1);
	return 0;
}

bool
parallel_ok(const fsm_b200_dfa *dfa, uint64_t len)
{
	/* chunk maps hold 16-bit states (DEAD16 reserved) */
	if (dfa->nstates > K1B_MAX_STATES) return false;
	uint64_t min_len = 4096;
	if (!small_table(dfa)) min_len = 1024ull * dfa->nstates;      /* at least a few chunks of 128 x T bytes */
	if (const char *e = getenv("FSM_B200_STREAM_MIN")) {
		const long v = atol(e);
		if (v >= 0) min_len = (uint64_t) v;
	}
	return len >= min_len && len >= 2 * PREFIX_W_MAX;
}

/* Serial form: a K1 batch of one (correct for any table size). */
int
stream_serial(const fsm_b200_dfa *dfa, const uint8_t *d_buf, uint64_t len, cudaStream_t stream,
	StreamScratch *ss, fsm_b200_result *out)
{
	if (ss->arena.cap < 4096) {
		if (ss->arena.base) cudaFree(ss->arena.base);
		ss->arena.base = nullptr; ss->arena.cap = 0;
		void *p = nullptr;
		FSMB_CUDA(cudaMalloc(&p, 1 << 20), return -1);
		ss->arena.base = static_cast<uint8_t *>(p);
		ss->arena.cap = 1 << 20;
	}
	fsm_b200_result *d_rec = reinterpret_cast<fsm_b200_result *>(ss->arena.base);
	if (k1_launch(dfa, d_buf, nullptr, len, len, 1, d_rec, stream, K1_LANE) != 0) return -1;
	FSMB_CUDA(cudaMemcpyAsync(out, d_rec, sizeof *out, cudaMemcpyDeviceToHost, stream), return -1);
	FSMB_CUDA(cudaStreamSynchronize(stream), return -1);
	return 0;
}

/* ---- one long input on an automaton with eager outputs ----------------------------------------
 * The chunk maps (above) give the exit state of every chunk for every entry state, but not the ids
 * fired on the way.  Second pass: follow the maps from the start state to get the TRUE entry state of
 * every chunk (one thread, nchunks dependent steps), then run the lines kernel over the chunks as
 * "lines" that start in those states and OR the per-chunk id sets.  The bytes are read twice; no walk
 * runs on a single lane. */
__global__ void
k1b_true_path_kernel(const StreamArgs a, uint32_t start, uint64_t *off, uint32_t *entry)
{
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	uint32_t st = start;
	uint32_t c = 0;
	for (; c < a.nchunks; c++) {
		off[c] = (uint64_t) c * a.C;
		entry[c] = st;
		const uint16_t nx = chunk_next(a, c, st, nullptr, nullptr);
		if (nx == DEAD16) { c++; break; }       /* chunk c's own walk stops at the missing edge */
		st = nx;
	}
	const uint64_t stop = min(a.len, (uint64_t) c * a.C);
	for (; c <= a.nchunks; c++) {               /* nothing after the chunk that died: empty lines */
		off[c] = stop;
		if (c < a.nchunks) entry[c] = start;
	}
}

template <int W>
__global__ void
k1b_or_masks_kernel(const uint64_t *masks, uint32_t n, uint64_t *out)
{
	__shared__ uint64_t part[W][256];
	uint64_t acc[W];
#pragma unroll
	for (int w = 0; w < W; w++) acc[w] = 0;
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
#pragma unroll
		for (int w = 0; w < W; w++) acc[w] |= masks[(size_t) i * W + w];
	}
#pragma unroll
	for (int w = 0; w < W; w++) part[w][threadIdx.x] = acc[w];
	__syncthreads();
	if (threadIdx.x < W) {
		uint64_t v = 0;
		for (uint32_t t = 0; t < blockDim.x; t++) v |= part[threadIdx.x][t];
		out[threadIdx.x] = v;
	}
}

} // namespace

namespace fsmb200 {
void
stream_scratch_free(fsm_b200_dfa *dfa)
{
	StreamScratch *ss = static_cast<StreamScratch *>(dfa->stream_scratch);
	if (ss == nullptr) return;
	cudaSetDevice(dfa->device);
	if (ss->stream) { cudaStreamSynchronize(ss->stream); cudaStreamDestroy(ss->stream); }
	cudaFree(ss->arena.base);
	cudaFree(ss->d_in);
	cudaFree(ss->d_eager);
	cudaFree(ss->d_dense);
	cudaFree(ss->rep_arena.base);
	if (ss->h_pinned) cudaFreeHost(ss->h_pinned);
	delete ss;
	dfa->stream_scratch = nullptr;
}

bool
k1b_stream_eager_ok(const fsm_b200_dfa *dfa, uint64_t len)
{
	return dfa->eager_nbits != 0 && dfa->eager_words >= 1 && dfa->eager_words <= 4 && k1_lines_eligible(dfa) &&
	    dfa->d_lperm != nullptr && parallel_ok(dfa, len) && getenv("FSM_B200_NO_EAGER_STREAM") == nullptr;
}

int
k1b_exec_stream_eager(const fsm_b200_dfa *dfa, const uint8_t *d_buf, uint64_t len,
	fsm_b200_result *h_rec, uint64_t *h_masks, cudaStream_t stream)
{
	StreamScratch *ss = ss_get(dfa);
	if (ss == nullptr) { errno = ENOMEM; return -1; }
	std::lock_guard<std::mutex> g(ss->mu);
	std::vector<StreamOut> m;
	StreamArgs a;
	if (stream_map(dfa, d_buf, len, stream, ss, m, &a) != 0) return -1;
	const StreamOut &r = m[dfa->start];
	h_rec->end = r.state;
	if (r.died) { h_rec->ret = 0; h_rec->consumed = r.dead_off; }
	else { h_rec->ret = dfa->h_is_end[r.state] ? 1 : 0; h_rec->consumed = len; }

	const uint32_t W = dfa->eager_words, nch = a.nchunks;
	const size_t need = (size_t) (nch + 1) * 8 + (size_t) nch * 4 + 256 + (size_t) nch * sizeof(fsm_b200_result) +
	    (size_t) nch * W * 8 + 256 + 64;
	if (ss->eager_cap < need) {
		if (ss->d_eager) cudaFree(ss->d_eager);
		ss->d_eager = nullptr; ss->eager_cap = 0;
		void *p = nullptr;
		FSMB_CUDA(cudaMalloc(&p, need + need / 4), return -1);
		ss->d_eager = static_cast<uint8_t *>(p);
		ss->eager_cap = need + need / 4;
	}
	Arena ar; ar.base = ss->d_eager; ar.cap = ss->eager_cap; ar.used = 0;
	uint64_t *d_off = ar.take<uint64_t>(nch + 1);
	uint32_t *d_entry = ar.take<uint32_t>(nch);
	fsm_b200_result *d_rec = ar.take<fsm_b200_result>(nch);
	uint64_t *d_masks = ar.take<uint64_t>((size_t) nch * W);
	uint64_t *d_final = ar.take<uint64_t>(8);
	k1b_true_path_kernel<<<1, 32, 0, stream>>>(a, dfa->start, d_off, d_entry);
	count_launch();
	if (k1_lines_launch(dfa, d_buf, d_off, 0, 0, nch, d_rec, d_masks, stream, d_entry) != 0) return -1;
	switch (W) {
	case 1: k1b_or_masks_kernel<1><<<1, 256, 0, stream>>>(d_masks, nch, d_final); break;
	case 2: k1b_or_masks_kernel<2><<<1, 256, 0, stream>>>(d_masks, nch, d_final); break;
	case 3: k1b_or_masks_kernel<3><<<1, 256, 0, stream>>>(d_masks, nch, d_final); break;
	default: k1b_or_masks_kernel<4><<<1, 256, 0, stream>>>(d_masks, nch, d_final); break;
	}
	count_launch();
	FSMB_CUDA(cudaGetLastError(), return -1);
	FSMB_CUDA(cudaMemcpyAsync(h_masks, d_final, W * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream), return -1);
	FSMB_CUDA(cudaStreamSynchronize(stream), return -1);
	return 0;
}
}

/* exec_stream on device memory; the caller holds ss->mu (scratch arena, d_in and stream are shared
 * by every call on this DFA). */
static int
exec_stream_dev_locked(const fsm_b200_dfa *dfa, const uint8_t *d_buf, uint64_t len,
	struct fsm_b200_result *out, cudaStream_t st, StreamScratch *ss)
{
	if (len == 0) {
		out->ret = dfa->h_is_end[dfa->start] ? 1 : 0;
		out->end = dfa->start;
		out->consumed = 0;
		return 0;
	}
	if (!parallel_ok(dfa, len) && !rep_eligible(dfa)) {      /* the fused small-automaton form takes any length */
		return stream_serial(dfa, d_buf, len, st, ss, out);
	}
	std::vector<StreamOut> m;
	if (stream_map(dfa, d_buf, len, st, ss, m) != 0) return -1;
	const StreamOut &r = m[dfa->start];
	out->end = r.state;
	if (r.died) {
		out->ret = 0;
		out->consumed = r.dead_off;
	} else {
		out->ret = dfa->h_is_end[r.state] ? 1 : 0;
		out->consumed = len;
	}
	return 0;
}

extern "C" int
fsm_b200_exec_stream_dev(const fsm_b200_dfa *dfa, const uint8_t *d_buf, uint64_t len,
	struct fsm_b200_result *out, void *stream)
{
	if (dfa == nullptr || out == nullptr || (len > 0 && d_buf == nullptr)) {
		set_error("exec_stream_dev: bad argument");
		errno = EINVAL;
		return -1;
	}
	FSMB_CUDA(cudaSetDevice(dfa->device), return -1);
	StreamScratch *ss = ss_get(dfa);
	if (ss == nullptr) { errno = ENOMEM; return -1; }
	std::lock_guard<std::mutex> g(ss->mu);
	return exec_stream_dev_locked(dfa, d_buf, len, out, static_cast<cudaStream_t>(stream), ss);
}

extern "C" int
fsm_b200_exec_stream_map_dev(const fsm_b200_dfa *dfa, const uint8_t *d_buf, uint64_t len,
	uint32_t *map_state, uint64_t *map_dead, uint32_t *map_dead_state, void *stream)
{
	if (dfa == nullptr || map_state == nullptr || map_dead == nullptr || map_dead_state == nullptr ||
	    (len > 0 && d_buf == nullptr)) {
		set_error("exec_stream_map_dev: bad argument");
		errno = EINVAL;
		return -1;
	}
	if (dfa->nstates > K1B_MAX_STATES) {
		set_error("exec_stream_map_dev: needs a table with at most %u states", K1B_MAX_STATES);
		errno = ENOTSUP;
		return -1;
	}
	FSMB_CUDA(cudaSetDevice(dfa->device), return -1);
	StreamScratch *ss = ss_get(dfa);
	if (ss == nullptr) { errno = ENOMEM; return -1; }
	std::lock_guard<std::mutex> g(ss->mu);
	const uint32_t T = dfa->nstates;
	if (len == 0) {
		for (uint32_t s = 0; s < T; s++) { map_state[s] = s; map_dead[s] = UINT64_MAX; map_dead_state[s] = UINT32_MAX; }
	} else {
		std::vector<StreamOut> m;
		if (stream_map(dfa, d_buf, len, static_cast<cudaStream_t>(stream), ss, m) != 0) return -1;
		for (uint32_t s = 0; s < T; s++) {
			if (m[s].died) {
				map_state[s] = dfa->dead; map_dead[s] = m[s].dead_off; map_dead_state[s] = m[s].state;
			} else {
				map_state[s] = m[s].state; map_dead[s] = UINT64_MAX; map_dead_state[s] = UINT32_MAX;
			}
		}
	}
	if (!dfa->complete) {                 /* the dead row maps to itself, dying at offset 0 */
		map_state[dfa->dead] = dfa->dead; map_dead[dfa->dead] = 0; map_dead_state[dfa->dead] = dfa->dead;
	}
	return 0;
}

/* The shard map left ON THE DEVICE, nothing waited for: the [nstates] records are written by work queued on
 * `stream`, so that a collective queued behind it on the same stream (the all-gather of the ranks' maps)
 * needs no host round trip. */
extern "C" int
fsm_b200_exec_stream_map_dev_async(const fsm_b200_dfa *dfa, const uint8_t *d_buf, uint64_t len,
	struct fsm_b200_stream_map_entry *d_map, void *stream)
{
	static_assert(sizeof(struct fsm_b200_stream_map_entry) == sizeof(StreamOut), "record layout");
	if (dfa == nullptr || d_map == nullptr || (len > 0 && d_buf == nullptr)) {
		set_error("exec_stream_map_dev_async: bad argument");
		errno = EINVAL;
		return -1;
	}
	if (dfa->nstates > K1B_MAX_STATES) {
		set_error("exec_stream_map_dev_async: needs a table with at most %u states", K1B_MAX_STATES);
		errno = ENOTSUP;
		return -1;
	}
	FSMB_CUDA(cudaSetDevice(dfa->device), return -1);
	StreamScratch *ss = ss_get(dfa);
	if (ss == nullptr) { errno = ENOMEM; return -1; }
	std::lock_guard<std::mutex> g(ss->mu);
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	const uint32_t T = dfa->nstates;
	if (len == 0 || (!parallel_ok(dfa, len) && !rep_eligible(dfa))) {
		/* nothing to cut into chunks: the synchronous form, then one copy (rare: empty or tiny shards) */
		std::vector<StreamOut> m(T);
		if (len == 0) {
			for (uint32_t s = 0; s < T; s++) { m[s].state = s; m[s].died = 0; m[s].dead_off = UINT64_MAX; }
		} else {
			std::vector<uint8_t> h((size_t) len);
			FSMB_CUDA(cudaMemcpyAsync(h.data(), d_buf, (size_t) len, cudaMemcpyDeviceToHost, st), return -1);
			FSMB_CUDA(cudaStreamSynchronize(st), return -1);
			for (uint32_t s = 0; s < T; s++) {
				uint32_t cur = s; uint64_t k = 0; bool died = false;
				for (; k < len; k++) {
					const uint32_t nx = dfa->h_table32[(size_t) cur * 256 + h[k]];
					if (nx == NO_EDGE) { died = true; break; }
					cur = nx;
				}
				m[s].state = cur; m[s].died = died ? 1u : 0u; m[s].dead_off = died ? k : UINT64_MAX;
			}
		}
		FSMB_CUDA(cudaMemcpyAsync(d_map, m.data(), T * sizeof(StreamOut), cudaMemcpyHostToDevice, st), return -1);
		FSMB_CUDA(cudaStreamSynchronize(st), return -1);
		return 0;
	}
	std::vector<StreamOut> unused;
	return stream_map(dfa, d_buf, len, st, ss, unused, nullptr, reinterpret_cast<StreamOut *>(d_map));
}

extern "C" int
fsm_b200_exec_stream_host(const fsm_b200_dfa *dfa, const uint8_t *buf, uint64_t len,
	struct fsm_b200_result *out)
{
	if (dfa == nullptr || out == nullptr || (len > 0 && buf == nullptr)) {
		set_error("exec_stream_host: bad argument");
		errno = EINVAL;
		return -1;
	}
	FSMB_CUDA(cudaSetDevice(dfa->device), return -1);
	StreamScratch *ss = ss_get(dfa);
	if (ss == nullptr) { errno = ENOMEM; return -1; }
	/* ss->mu is held across copy + scan + read-back: two threads running fsm_exec on the SAME fsm
	 * share the cached DFA and with it d_in and the stream (the lock used to be dropped
	 * between the H2D copy and the scan) */
	std::lock_guard<std::mutex> g(ss->mu);
	if (ss->stream == nullptr) {
		FSMB_CUDA(cudaStreamCreateWithFlags(&ss->stream, cudaStreamNonBlocking), return -1);
	}
	if (ss->in_cap < len + 64) {
		if (ss->d_in) cudaFree(ss->d_in);
		ss->d_in = nullptr; ss->in_cap = 0;
		void *p = nullptr;
		FSMB_CUDA(cudaMalloc(&p, len + len / 8 + 4096), return -1);
		ss->d_in = static_cast<uint8_t *>(p);
		ss->in_cap = len + len / 8 + 4096;
	}
	if (len > 0) {
		FSMB_CUDA(cudaMemcpyAsync(ss->d_in, buf, len, cudaMemcpyHostToDevice, ss->stream), return -1);
	}
	return exec_stream_dev_locked(dfa, ss->d_in, len, out, ss->stream, ss);
}
