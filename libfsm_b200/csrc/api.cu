/*
 * api.cu -- C-ABI glue: error state, batch entry points (host and device pointers).
 *
 * fsm_b200_exec_batch_host is the end-to-end path a relinked re(1)/fsm(1) reaches through
 * the shim's fsm_exec / fsm_exec_batch: host buffers in, host records out, with the
 * host<->device copies pipelined in chunks over two streams.
 */
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "common.h"
#include "k1_exec_batch.h"

namespace fsmb200 {

static thread_local char tl_error[512] = "";
static thread_local uint64_t tl_launches = 0;

void
set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(tl_error, sizeof tl_error, fmt, ap);
	va_end(ap);
}

void
count_launch(uint64_t n)
{
	tl_launches += n;
}


/* ---- pinned host block pool (common.h) ---- */
namespace {
struct PoolBlock { void *p; size_t cap; };
std::mutex g_pool_mu;
std::vector<PoolBlock> g_pool_free;
size_t g_pool_free_bytes = 0;
constexpr size_t POOL_MAX_BLOCKS = 8, POOL_MAX_BYTES = 1ull << 30;
}

void *
host_pool_get(size_t bytes, size_t *cap)
{
	if (bytes == 0) bytes = 1;
	{
		std::lock_guard<std::mutex> g(g_pool_mu);
		size_t best = g_pool_free.size();
		for (size_t i = 0; i < g_pool_free.size(); i++) {
			if (g_pool_free[i].cap >= bytes && (best == g_pool_free.size() || g_pool_free[i].cap < g_pool_free[best].cap)) best = i;
		}
		if (best != g_pool_free.size() && g_pool_free[best].cap <= 4 * bytes + (1u << 20)) {
			const PoolBlock b = g_pool_free[best];
			g_pool_free.erase(g_pool_free.begin() + (long) best);
			g_pool_free_bytes -= b.cap;
			*cap = b.cap;
			return b.p;
		}
	}
	/* 12 % headroom, 1 MiB granules: the next result of about this size fits the same block */
	const size_t want = (bytes + (bytes >> 3) + 0xFFFFFu) & ~(size_t) 0xFFFFFu;
	void *p = nullptr;
	if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	*cap = want;
	return p;
}

void
host_pool_put(void *p, size_t cap)
{
	if (p == nullptr) return;
	{
		std::lock_guard<std::mutex> g(g_pool_mu);
		if (g_pool_free.size() < POOL_MAX_BLOCKS && g_pool_free_bytes + cap <= POOL_MAX_BYTES) {
			g_pool_free.push_back(PoolBlock{ p, cap });
			g_pool_free_bytes += cap;
			return;
		}
	}
	cudaFreeHost(p);
}

/* Per-DFA scratch for the _host entry points: two slots (double buffering). */
struct Slot {
	cudaStream_t stream = nullptr;
	uint8_t *d_in = nullptr;      size_t in_cap = 0;
	uint64_t *d_off = nullptr;    size_t off_cap = 0;     /* entries */
	fsm_b200_result *d_out = nullptr; size_t out_cap = 0; /* entries */
	bool busy = false;
};

struct Scratch {
	std::mutex mu;
	Slot slot[2];
};

static Scratch *
scratch_get(const fsm_b200_dfa *cdfa)
{
	fsm_b200_dfa *dfa = const_cast<fsm_b200_dfa *>(cdfa);
	static std::mutex create_mu;
	std::lock_guard<std::mutex> g(create_mu);
	if (dfa->scratch == nullptr) {
		dfa->scratch = new (std::nothrow) Scratch();
	}
	return static_cast<Scratch *>(dfa->scratch);
}

void
scratch_free(fsm_b200_dfa *dfa)
{
	Scratch *sc = static_cast<Scratch *>(dfa->scratch);
	if (sc == nullptr) return;
	cudaSetDevice(dfa->device);
	for (Slot &s : sc->slot) {
		if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
		cudaFree(s.d_in); cudaFree(s.d_off); cudaFree(s.d_out);
	}
	delete sc;
	dfa->scratch = nullptr;
}

template <typename T>
static bool
grow(T **p, size_t *cap, size_t want)
{
	if (*cap >= want) return true;
	if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
	size_t ncap = want + want / 4 + 256;
	void *q = nullptr;
	if (cudaMalloc(&q, ncap * sizeof(T)) != cudaSuccess) {
		cudaGetLastError();
		ncap = want;
		if (cudaMalloc(&q, ncap * sizeof(T)) != cudaSuccess) {
			cudaGetLastError();
			return false;
		}
	}
	*p = static_cast<T *>(q);
	*cap = ncap;
	return true;
}

} // namespace fsmb200

using namespace fsmb200;

extern "C" int
fsm_b200_abi_version(void)
{
	return FSM_B200_ABI_VERSION;
}

extern "C" int
fsm_b200_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	int usable = 0;
	for (int i = 0; i < n; i++) {
		int major = 0;
		if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) {
			usable++;
		}
	}
	return usable;
}

extern "C" const char *
fsm_b200_last_error(void)
{
	return tl_error;
}

extern "C" uint64_t
fsm_b200_launch_count(int reset)
{
	const uint64_t v = tl_launches;
	if (reset) tl_launches = 0;
	return v;
}

extern "C" int
fsm_b200_exec_batch_dev(const fsm_b200_dfa *dfa,
	const uint8_t *d_base, const uint64_t *d_offsets, uint64_t stride, uint64_t len,
	size_t n, struct fsm_b200_result *d_out, void *stream)
{
	if (dfa == nullptr || (n > 0 && (d_base == nullptr || d_out == nullptr))) {
		set_error("exec_batch_dev: bad argument");
		errno = EINVAL;
		return -1;
	}
	if (d_offsets == nullptr && stride < len) {
		set_error("exec_batch_dev: stride < len");
		errno = EINVAL;
		return -1;
	}
	FSMB_CUDA(cudaSetDevice(dfa->device), return -1);
	return k1_launch(dfa, d_base, d_offsets, stride, len, n, d_out, static_cast<cudaStream_t>(stream), K1_AUTO);
}

extern "C" int
fsm_b200_exec_batch_host(const fsm_b200_dfa *dfa,
	const uint8_t *base, const uint64_t *offsets, size_t n,
	struct fsm_b200_result *out)
{
	if (dfa == nullptr || (n > 0 && (base == nullptr || offsets == nullptr || out == nullptr))) {
		set_error("exec_batch_host: bad argument");
		errno = EINVAL;
		return -1;
	}
	if (n == 0) return 0;
	for (size_t i = 0; i < n; i++) {
		if (offsets[i + 1] < offsets[i]) {
			set_error("exec_batch_host: offsets not monotone at %zu", i);
			errno = EINVAL;
			return -1;
		}
	}
	FSMB_CUDA(cudaSetDevice(dfa->device), return -1);
	Scratch *sc = scratch_get(dfa);
	if (sc == nullptr) {
		errno = ENOMEM;
		return -1;
	}
	std::lock_guard<std::mutex> guard(sc->mu);

	/* fixed stride?  (lets the kernel skip the offsets array and use the TMA-tiled variant) */
	const uint64_t len0 = offsets[1] - offsets[0];
	bool fixed = true;
	for (size_t i = 1; i < n && fixed; i++) {
		fixed = (offsets[i + 1] - offsets[i]) == len0;
	}

	size_t chunk_bytes = 64u << 20;
	if (const char *e = getenv("FSM_B200_HOST_CHUNK_MB")) {
		long v = atol(e);
		if (v >= 1 && v <= 4096) chunk_bytes = (size_t) v << 20;
	}

	int rc = 0;
	size_t i0 = 0;
	int which = 0;
	while (i0 < n && rc == 0) {
		/* chunk = [i0, i1): at least one input, at most ~chunk_bytes of bytes */
		size_t i1 = i0 + 1;
		if (fixed) {
			size_t per = len0 > 0 ? chunk_bytes / len0 : (size_t) (8u << 20);
			if (per < 1) per = 1;
			if (per > (8u << 20)) per = 8u << 20;
			i1 = (n - i0 < per) ? n : i0 + per;
		} else {
			/* binary search the offsets for the chunk end */
			const uint64_t limit = offsets[i0] + chunk_bytes;
			size_t lo = i0 + 1, hi = n;
			while (lo < hi) {
				size_t mid = lo + (hi - lo + 1) / 2;
				if (offsets[mid] <= limit) lo = mid; else hi = mid - 1;
			}
			i1 = lo;
			if (i1 - i0 > (8u << 20)) i1 = i0 + (8u << 20);    /* bound records per chunk */
		}
		const size_t cn = i1 - i0;
		const uint64_t lo_b = offsets[i0], hi_b = offsets[i1];
		const size_t nbytes = (size_t) (hi_b - lo_b);

		Slot &s = sc->slot[which];
		which ^= 1;
		if (s.stream == nullptr) {
			FSMB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), { rc = -1; break; });
		}
		if (s.busy) {
			FSMB_CUDA(cudaStreamSynchronize(s.stream), { rc = -1; break; });
			s.busy = false;
		}
		if (!grow(&s.d_in, &s.in_cap, nbytes + 512) || !grow(&s.d_out, &s.out_cap, cn) ||
		    (!fixed && !grow(&s.d_off, &s.off_cap, cn + 1))) {
			set_error("exec_batch_host: out of device memory");
			errno = ENOMEM;
			rc = -1;
			break;
		}
		/* keep each input's alignment modulo 256 so aligned batches stay aligned */
		uint8_t *d_chunk = s.d_in + (fixed ? 0 : (size_t) (lo_b & 255u));
		if (nbytes > 0) {
			FSMB_CUDA(cudaMemcpyAsync(d_chunk, base + lo_b, nbytes, cudaMemcpyHostToDevice, s.stream),
			    { rc = -1; break; });
		}
		const uint8_t *d_base = d_chunk - lo_b;   /* so that d_base + offsets[i] is input i */
		if (fixed) {
			rc = k1_launch(dfa, d_chunk, nullptr, len0, len0, cn, s.d_out, s.stream, K1_AUTO);
		} else {
			FSMB_CUDA(cudaMemcpyAsync(s.d_off, offsets + i0, (cn + 1) * sizeof(uint64_t),
			    cudaMemcpyHostToDevice, s.stream), { rc = -1; break; });
			rc = k1_launch(dfa, d_base, s.d_off, 0, 0, cn, s.d_out, s.stream, K1_AUTO);
		}
		if (rc != 0) break;
		FSMB_CUDA(cudaMemcpyAsync(out + i0, s.d_out, cn * sizeof(fsm_b200_result),
		    cudaMemcpyDeviceToHost, s.stream), { rc = -1; break; });
		s.busy = true;
		i0 = i1;
	}
	for (Slot &s : sc->slot) {
		if (s.busy) {
			cudaError_t e = cudaStreamSynchronize(s.stream);
			s.busy = false;
			if (e != cudaSuccess && rc == 0) {
				set_error("exec_batch_host: %s", cudaGetErrorString(e));
				errno = EIO;
				rc = -1;
			}
		}
	}
	return rc;
}


/* ---- fused scan + gather over NVLink peer memory ------------------------------------------ */

extern "C" int
fsm_b200_exec_batch_dev_gather(const fsm_b200_dfa *dfa,
	const uint8_t *d_base, const uint64_t *d_offsets, uint64_t stride, uint64_t len, size_t n,
	struct fsm_b200_result *d_out, void *const *peer_outs, int npeers, int compact,
	void *sig_counter, void *const *sig_flags, uint32_t sig_value, void *stream)
{
	if (dfa == nullptr || (n > 0 && (d_base == nullptr || d_out == nullptr)) || (npeers > 0 && peer_outs == nullptr)) {
		set_error("exec_batch_dev_gather: bad argument");
		errno = EINVAL;
		return -1;
	}
	FSMB_CUDA(cudaSetDevice(dfa->device), return -1);
	return k1_launch(dfa, d_base, d_offsets, stride, len, n, d_out, static_cast<cudaStream_t>(stream), K1_AUTO,
	    reinterpret_cast<fsm_b200_result *const *>(peer_outs), npeers, compact,
	    static_cast<uint32_t *>(sig_counter), reinterpret_cast<uint32_t *const *>(sig_flags), sig_value);
}

/* Consumer side of the completion flags: one warp polls flags[0..n) (system scope) until every one has
 * reached `value`; bounded (about a second) so that a peer that died cannot hang the device. */
__global__ void
wait_flags_kernel(const volatile uint32_t *flags, uint32_t n, uint32_t value, uint32_t *timed_out)
{
	const uint32_t r = threadIdx.x;
	if (r >= n) return;
	const long long t0 = clock64();
	while ((int32_t) (flags[r] - value) < 0) {
		__nanosleep(200);
		if (clock64() - t0 > 2000000000ll) { if (timed_out != nullptr) atomicExch(timed_out, 1u); break; }
	}
	__threadfence_system();
}

extern "C" int
fsm_b200_wait_flags_dev(int device, const void *d_flags, uint32_t n, uint32_t value, void *d_timed_out, void *stream)
{
	if (d_flags == nullptr || n == 0 || n > 32) {
		set_error("wait_flags_dev: bad argument");
		errno = EINVAL;
		return -1;
	}
	FSMB_CUDA(cudaSetDevice(device), return -1);
	wait_flags_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const volatile uint32_t *>(d_flags), n, value,
	    static_cast<uint32_t *>(d_timed_out));
	count_launch();
	FSMB_CUDA(cudaGetLastError(), return -1);
	return 0;
}

extern "C" int
fsm_b200_dev_alloc(int device, size_t bytes, void **out)
{
	if (out == nullptr) { errno = EINVAL; return -1; }
	FSMB_CUDA(cudaSetDevice(device), return -1);
	FSMB_CUDA(cudaMalloc(out, bytes), return -1);
	return 0;
}

extern "C" int
fsm_b200_dev_free(int device, void *p)
{
	FSMB_CUDA(cudaSetDevice(device), return -1);
	FSMB_CUDA(cudaFree(p), return -1);
	return 0;
}

extern "C" int
fsm_b200_dev_zero(int device, void *p, size_t bytes)
{
	FSMB_CUDA(cudaSetDevice(device), return -1);
	FSMB_CUDA(cudaMemset(p, 0, bytes), return -1);
	return 0;
}

extern "C" int
fsm_b200_dev_read(int device, void *host_dst, const void *dev_src, size_t bytes)
{
	FSMB_CUDA(cudaSetDevice(device), return -1);
	FSMB_CUDA(cudaMemcpy(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost), return -1);
	return 0;
}

extern "C" int
fsm_b200_ipc_export(const void *dev_ptr, void *handle64)
{
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	cudaIpcMemHandle_t h;
	FSMB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)), return -1);
	memcpy(handle64, &h, 64);
	return 0;
}

extern "C" int
fsm_b200_ipc_open(int device, const void *handle64, void **out)
{
	cudaIpcMemHandle_t h;
	if (out == nullptr) { errno = EINVAL; return -1; }
	memcpy(&h, handle64, 64);
	FSMB_CUDA(cudaSetDevice(device), return -1);
	FSMB_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess), return -1);
	return 0;
}

extern "C" int
fsm_b200_ipc_close(int device, void *p)
{
	FSMB_CUDA(cudaSetDevice(device), return -1);
	FSMB_CUDA(cudaIpcCloseMemHandle(p), return -1);
	return 0;
}
