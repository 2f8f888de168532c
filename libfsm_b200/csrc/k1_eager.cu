/*
 * k1_eager.cu -- batch execution that also reports EAGER OUTPUTS.
 *
 * The reference's fsm_exec calls the fsm_eager_output_cb callback for every eager-output id of
 * the start state and of every state it enters (src/libfsm/exec.c:55-83,126-130,140-144;
 * include/fsm/fsm.h:273-336), whether or not the input ends up matching.  Here one lane walks
 * one input -- as in K1's LANE kernel -- and ORs the per-state id mask (W 64-bit words, built
 * once by fsm_b200_dfa_compile) into a register accumulator: the answer is the SET of fired ids
 * as a bitset per input, next to the usual 16-byte record.
 *
 * Round 1 status: correctness-first.  The table blob and the masks are read from global memory
 * through the read-only path (they are a few KB to a few MB and stay L1/L2-resident); the
 * shared-memory staging, k-stride stepping and vectorised input loads of k1_exec_batch.cu are
 * not applied yet.  One extra dependent load per byte is the price of the feature on any design:
 * mask[state] can only be fetched once the state is known.
 */
#include <cstring>
#include <mutex>
#include <new>

#include "common.h"
#include "k1_exec_batch.h"

using namespace fsmb200;

namespace {

/* per-DFA scratch of the _host entry point */
struct EagerSlot {
	cudaStream_t stream = nullptr;
	void *d_in = nullptr, *d_off = nullptr, *d_out = nullptr, *d_masks = nullptr;
	size_t in_cap = 0, off_cap = 0, out_cap = 0, masks_cap = 0;
	bool busy = false;
};

/* per-DFA grow-only buffers: two slots so that chunk k+1's host->device copy overlaps chunk k's scan and
 * device->host copy (PCIe is full duplex; records + id bitsets are a fifth of config 3's traffic) */
struct EagerScratch {
	std::mutex mu;
	EagerSlot slot[2];
	static bool grow(void **p, size_t *cap, size_t want) {
		if (*cap >= want) return true;
		if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
		const size_t ncap = want + want / 4 + 4096;
		if (cudaMalloc(p, ncap) != cudaSuccess) { cudaGetLastError(); *p = nullptr; return false; }
		*cap = ncap;
		return true;
	}
};

std::mutex g_eager_scratch_mu;

EagerScratch *
eager_scratch_get(const fsm_b200_dfa *cdfa)
{
	fsm_b200_dfa *dfa = const_cast<fsm_b200_dfa *>(cdfa);
	std::lock_guard<std::mutex> g(g_eager_scratch_mu);
	if (dfa->eager_scratch == nullptr) dfa->eager_scratch = new (std::nothrow) EagerScratch();
	return static_cast<EagerScratch *>(dfa->eager_scratch);
}

struct EagerArgs {
	const uint8_t *blob;          /* table rows | is_end | class LUT */
	const uint64_t *masks;        /* [ntable][W] */
	const uint8_t *base;
	const uint64_t *offsets;      /* nullptr: fixed stride */
	uint64_t stride, len, n;
	fsm_b200_result *out;
	uint64_t *out_masks;          /* [n][W] */
	uint32_t pitch, entry_bytes, nclasses, is_end_off, cls_off, start, dead;
};

template <int W>
__global__ void __launch_bounds__(256)
k1_eager_kernel(const EagerArgs a)
{
	const uint8_t *is_end = a.blob + a.is_end_off;
	const uint8_t *cls = a.blob + a.cls_off;
	const uint64_t nthreads = (uint64_t) gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += nthreads) {
		uint64_t beg, len;
		if (a.offsets != nullptr) { beg = a.offsets[i]; len = a.offsets[i + 1] - beg; }
		else { beg = i * a.stride; len = a.len; }
		const uint8_t *p = a.base + beg;
		uint32_t st = a.start;
		uint64_t acc[W];
#pragma unroll
		for (int w = 0; w < W; w++) acc[w] = __ldg(a.masks + (size_t) st * W + w);      /* exec.c:126-130 */
		uint64_t pos = 0;
		bool died = false;
		for (; pos < len; pos++) {
			uint32_t col = __ldg(p + pos);
			if (a.nclasses != 0) col = __ldg(cls + col);
			const uint8_t *row = a.blob + (size_t) st * a.pitch;
			uint32_t nx;
			if (a.entry_bytes == 1) nx = __ldg(row + col);
			else if (a.entry_bytes == 2) nx = __ldg(reinterpret_cast<const uint16_t *>(row) + col);
			else nx = __ldg(reinterpret_cast<const uint32_t *>(row) + col);
			if (nx == a.dead) { died = true; break; }       /* exec.c:133-138: no edge, stop reading */
			st = nx;
#pragma unroll
			for (int w = 0; w < W; w++) acc[w] |= __ldg(a.masks + (size_t) st * W + w);  /* exec.c:140-144 */
		}
		uint4 v;
		v.x = (uint32_t) ((!died && __ldg(is_end + st)) ? 1 : 0);
		v.y = st;
		v.z = (uint32_t) pos;
		v.w = (uint32_t) (pos >> 32);
		*reinterpret_cast<uint4 *>(a.out + i) = v;
#pragma unroll
		for (int w = 0; w < W; w++) a.out_masks[i * W + w] = acc[w];
	}
}

int
launch_eager(const fsm_b200_dfa *dfa, const uint8_t *d_base, const uint64_t *d_offsets, uint64_t stride, uint64_t len,
	size_t n, fsm_b200_result *d_out, uint64_t *d_masks, cudaStream_t stream)
{
	if (n == 0) return 0;
	if (k1_lines_eligible(dfa)) {          /* table fits shared memory: the tuned kernel (k1_lines.cu) */
		return k1_lines_launch(dfa, d_base, d_offsets, stride, len, n, d_out, d_masks, stream);
	}
	EagerArgs a;
	memset(&a, 0, sizeof a);
	a.blob = static_cast<const uint8_t *>(dfa->d_blob);
	a.masks = dfa->d_eager_masks;
	a.base = d_base; a.offsets = d_offsets; a.stride = stride; a.len = len; a.n = n;
	a.out = d_out; a.out_masks = d_masks;
	a.pitch = dfa->pitch; a.entry_bytes = dfa->entry_bytes; a.nclasses = dfa->nclasses;
	a.is_end_off = dfa->is_end_off; a.cls_off = dfa->cls_off; a.start = dfa->start; a.dead = dfa->dead;
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dfa->device);
	const uint64_t want = (n + 255) / 256;
	const unsigned blocks = (unsigned) (want < (uint64_t) sms * 8 ? want : (uint64_t) sms * 8);    /* grid-stride above that */
	switch (dfa->eager_words) {
	case 1: k1_eager_kernel<1><<<blocks, 256, 0, stream>>>(a); break;
	case 2: k1_eager_kernel<2><<<blocks, 256, 0, stream>>>(a); break;
	case 3: k1_eager_kernel<3><<<blocks, 256, 0, stream>>>(a); break;
	case 4: k1_eager_kernel<4><<<blocks, 256, 0, stream>>>(a); break;
	default:
		set_error("exec_batch_eager: %u mask words not supported", dfa->eager_words);
		errno = ENOTSUP;
		return -1;
	}
	count_launch();
	FSMB_CUDA(cudaGetLastError(), return -1);
	return 0;
}

} // namespace

extern "C" int
fsm_b200_exec_batch_eager_dev(const fsm_b200_dfa *dfa,
	const uint8_t *d_base, const uint64_t *d_offsets, uint64_t stride, uint64_t len,
	size_t n, struct fsm_b200_result *d_out, uint64_t *d_masks, void *stream)
{
	if (dfa == nullptr || (n > 0 && (d_base == nullptr || d_out == nullptr || d_masks == nullptr)) ||
	    (d_offsets == nullptr && stride < len)) {
		set_error("exec_batch_eager_dev: bad argument");
		errno = EINVAL;
		return -1;
	}
	if (dfa->eager_nbits == 0) {
		set_error("exec_batch_eager_dev: this DFA has no eager outputs (use fsm_b200_exec_batch_dev)");
		errno = EINVAL;
		return -1;
	}
	FSMB_CUDA(cudaSetDevice(dfa->device), return -1);
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	if (n == 1) {
		/* one (possibly long) input: K1b's chunked form when it is long enough, never one lane */
		uint64_t beg = 0, ilen = len;
		if (d_offsets != nullptr) {
			uint64_t h[2];
			FSMB_CUDA(cudaMemcpyAsync(h, d_offsets, sizeof h, cudaMemcpyDeviceToHost, st), return -1);
			FSMB_CUDA(cudaStreamSynchronize(st), return -1);
			beg = h[0]; ilen = h[1] - h[0];
		}
		if (k1b_stream_eager_ok(dfa, ilen)) {
			fsm_b200_result rec;
			uint64_t hm[4] = { 0, 0, 0, 0 };
			if (k1b_exec_stream_eager(dfa, d_base + beg, ilen, &rec, hm, st) != 0) return -1;
			FSMB_CUDA(cudaMemcpyAsync(d_out, &rec, sizeof rec, cudaMemcpyHostToDevice, st), return -1);
			FSMB_CUDA(cudaMemcpyAsync(d_masks, hm, dfa->eager_words * sizeof(uint64_t), cudaMemcpyHostToDevice, st), return -1);
			FSMB_CUDA(cudaStreamSynchronize(st), return -1);     /* rec / hm live on this stack frame */
			return 0;
		}
	}
	return launch_eager(dfa, d_base, d_offsets, stride, len, n, d_out, d_masks, st);
}

extern "C" int
fsm_b200_exec_batch_eager_host(const fsm_b200_dfa *dfa,
	const uint8_t *base, const uint64_t *offsets, size_t n,
	struct fsm_b200_result *out, uint64_t *masks)
{
	if (dfa == nullptr || (n > 0 && (base == nullptr || offsets == nullptr || out == nullptr || masks == nullptr))) {
		set_error("exec_batch_eager_host: bad argument");
		errno = EINVAL;
		return -1;
	}
	if (dfa->eager_nbits == 0) {
		set_error("exec_batch_eager_host: this DFA has no eager outputs (use fsm_b200_exec_batch_host)");
		errno = EINVAL;
		return -1;
	}
	if (n == 0) return 0;
	for (size_t i = 0; i < n; i++) {
		if (offsets[i + 1] < offsets[i]) {
			set_error("exec_batch_eager_host: offsets not monotone at %zu", i);
			errno = EINVAL;
			return -1;
		}
	}
	FSMB_CUDA(cudaSetDevice(dfa->device), return -1);
	const size_t W = dfa->eager_words;
	/* grow-only buffers and streams per DFA, under a mutex (the shim's fsm_exec comes here for every
	 * input of an automaton with eager outputs: no cudaMalloc / cudaFree / stream creation per call) */
	EagerScratch *sc = eager_scratch_get(dfa);
	if (sc == nullptr) { errno = ENOMEM; return -1; }
	std::lock_guard<std::mutex> guard(sc->mu);

	size_t chunk_bytes = 64u << 20;
	if (const char *e = getenv("FSM_B200_HOST_CHUNK_MB")) {
		long v = atol(e);
		if (v >= 1 && v <= 4096) chunk_bytes = (size_t) v << 20;
	}
	if (n == 1 && k1b_stream_eager_ok(dfa, offsets[1] - offsets[0])) {
		/* one long input (the shim's fsm_exec on an automaton with eager outputs): K1b's chunked form */
		EagerSlot &s = sc->slot[0];
		const uint64_t lo_b = offsets[0], nbytes = offsets[1] - offsets[0];
		if (s.stream == nullptr) FSMB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), return -1);
		if (!EagerScratch::grow(&s.d_in, &s.in_cap, nbytes + 64)) {
			set_error("exec_batch_eager_host: out of device memory");
			errno = ENOMEM;
			return -1;
		}
		FSMB_CUDA(cudaMemcpyAsync(s.d_in, base + lo_b, nbytes, cudaMemcpyHostToDevice, s.stream), return -1);
		for (size_t w = 0; w < W; w++) masks[w] = 0;
		return k1b_exec_stream_eager(dfa, static_cast<const uint8_t *>(s.d_in), nbytes, out, masks, s.stream);
	}
	int rc = 0, which = 0;
	size_t i0 = 0;
	while (i0 < n && rc == 0) {
		/* chunk = lines [i0, i1): at least one, at most ~chunk_bytes of bytes and 4 M lines */
		const uint64_t limit = offsets[i0] + chunk_bytes;
		size_t lo = i0 + 1, hi = n;
		while (lo < hi) {
			const size_t mid = lo + (hi - lo + 1) / 2;
			if (offsets[mid] <= limit) lo = mid; else hi = mid - 1;
		}
		size_t i1 = lo;
		if (i1 - i0 > (4u << 20)) i1 = i0 + (4u << 20);
		const size_t cn = i1 - i0;
		const uint64_t lo_b = offsets[i0];
		const size_t nbytes = (size_t) (offsets[i1] - lo_b);

		EagerSlot &s = sc->slot[which];
		which ^= 1;
		if (s.stream == nullptr) {
			FSMB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), { rc = -1; break; });
		}
		if (s.busy) {
			FSMB_CUDA(cudaStreamSynchronize(s.stream), { rc = -1; break; });
			s.busy = false;
		}
		if (!EagerScratch::grow(&s.d_in, &s.in_cap, nbytes + 64) || !EagerScratch::grow(&s.d_off, &s.off_cap, (cn + 1) * sizeof(uint64_t)) ||
		    !EagerScratch::grow(&s.d_out, &s.out_cap, cn * sizeof(fsm_b200_result)) || !EagerScratch::grow(&s.d_masks, &s.masks_cap, cn * W * sizeof(uint64_t))) {
			set_error("exec_batch_eager_host: out of device memory");
			errno = ENOMEM;
			rc = -1;
			break;
		}
		/* keep the alignment of the caller's bytes modulo 32 (sector loads) */
		uint8_t *d_in = static_cast<uint8_t *>(s.d_in) + (lo_b & 31u);
		if (nbytes > 0) FSMB_CUDA(cudaMemcpyAsync(d_in, base + lo_b, nbytes, cudaMemcpyHostToDevice, s.stream), { rc = -1; break; });
		FSMB_CUDA(cudaMemcpyAsync(s.d_off, offsets + i0, (cn + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s.stream), { rc = -1; break; });
		if (launch_eager(dfa, d_in - lo_b, static_cast<const uint64_t *>(s.d_off), 0, 0, cn, static_cast<fsm_b200_result *>(s.d_out),
		    static_cast<uint64_t *>(s.d_masks), s.stream) != 0) {
			rc = -1;
			break;
		}
		FSMB_CUDA(cudaMemcpyAsync(out + i0, s.d_out, cn * sizeof(fsm_b200_result), cudaMemcpyDeviceToHost, s.stream), { rc = -1; break; });
		FSMB_CUDA(cudaMemcpyAsync(masks + i0 * W, s.d_masks, cn * W * sizeof(uint64_t), cudaMemcpyDeviceToHost, s.stream), { rc = -1; break; });
		s.busy = true;
		i0 = i1;
	}
	for (EagerSlot &s : sc->slot) {
		if (s.stream != nullptr && (s.busy || rc != 0)) {
			cudaError_t e = cudaStreamSynchronize(s.stream);
			s.busy = false;
			if (e != cudaSuccess && rc == 0) {
				set_error("exec_batch_eager_host: %s", cudaGetErrorString(e));
				errno = EIO;
				rc = -1;
			}
		}
	}
	return rc;
}

namespace fsmb200 {
void
eager_scratch_free(fsm_b200_dfa *dfa)
{
	EagerScratch *sc = static_cast<EagerScratch *>(dfa->eager_scratch);
	if (sc == nullptr) return;
	cudaSetDevice(dfa->device);
	for (EagerSlot &s : sc->slot) {
		if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
		cudaFree(s.d_in); cudaFree(s.d_off); cudaFree(s.d_out); cudaFree(s.d_masks);
	}
	delete sc;
	dfa->eager_scratch = nullptr;
}
}
