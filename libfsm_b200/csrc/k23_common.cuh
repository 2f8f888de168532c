/*
 * k23_common.cuh -- pieces shared by K2 (determinise) and K3 (minimise): stream-ordered
 * device buffers, an exclusive-scan primitive (no CUB/Thrust), byte-class analysis of label
 * sets, the group emitter and the owned output description.  Included inside an anonymous
 * namespace by each translation unit.
 */
#ifndef FSM_B200_K23_COMMON_CUH
#define FSM_B200_K23_COMMON_CUH

#include <algorithm>
#include <array>
#include <chrono>
#include <cstring>
#include <map>
#include <new>
#include <vector>

#include "common.h"

namespace {

using namespace fsmb200;

constexpr uint32_t NONE32 = 0xFFFFFFFFu;
constexpr uint64_t EMPTY64 = 0xFFFFFFFFFFFFFFFFull;

#define CK(expr) FSMB_CUDA(expr, return -1)
/* Synchronise AND pick up a failed launch: a launch that could not start (bad configuration, too
 * much shared memory) leaves a non-sticky error that only cudaGetLastError reports.  The entry points
 * clear the thread's error state first, so whatever is found here belongs to this call. */
#define CK_SYNC(st) do { CK(cudaStreamSynchronize(st)); CK(cudaGetLastError()); } while (0)

/* ------------------------------------------------------------------ device buffers ---- */

/* Stream-ordered allocations from the device's default memory pool (release threshold raised
 * in fsm_b200_determinise so memory is retained across calls: no cudaMalloc/cudaFree
 * synchronisation inside the frontier loop). */
template <typename T> struct DBuf {
	T *p = nullptr;
	size_t cap = 0;
	cudaStream_t owner = nullptr;
	~DBuf() { if (p) cudaFreeAsync(p, owner); }
	int reserve(size_t n, bool keep, cudaStream_t st) {
		if (n <= cap) return 0;
		owner = st;
		size_t ncap = std::max(n, cap + cap / 2 + 1024);
		T *q = nullptr;
		cudaError_t e = cudaMallocAsync(&q, ncap * sizeof(T), st);
		if (e != cudaSuccess) {
			cudaGetLastError();
			ncap = n;
			e = cudaMallocAsync(&q, ncap * sizeof(T), st);
			if (e != cudaSuccess) {
				set_error("determinise: cudaMalloc(%zu bytes) failed: %s", ncap * sizeof(T), cudaGetErrorString(e));
				errno = ENOMEM;
				return -1;
			}
		}
		if (keep && p && cap) {
			if (cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
				cudaFreeAsync(q, st);
				set_error("determinise: device copy failed");
				errno = EIO;
				return -1;
			}
		}
		if (p) cudaFreeAsync(p, st);
		p = q; cap = ncap;
		return 0;
	}
};

/* ------------------------------------------------------------------ exclusive scan ---- */

constexpr int SCAN_T = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_T * SCAN_ITEMS;

template <typename TIn>
__global__ void
scan_tile_kernel(const TIn *in, uint64_t *out, uint64_t *tile_sums, uint64_t n)
{
	__shared__ uint64_t warp_sums[SCAN_T / 32];
	const uint64_t base = (uint64_t) blockIdx.x * SCAN_TILE + (uint64_t) threadIdx.x * SCAN_ITEMS;
	uint64_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; i++) {
		v[i] = (base + i < n) ? (uint64_t) in[base + i] : 0;
		sum += v[i];
	}
	uint64_t incl = sum;
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
		if (lane >= (unsigned) d) incl += t;
	}
	if (lane == 31) warp_sums[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		uint64_t w = lane < SCAN_T / 32 ? warp_sums[lane] : 0;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, w, d);
			if (lane >= (unsigned) d) w += t;
		}
		if (lane < SCAN_T / 32) warp_sums[lane] = w;
	}
	__syncthreads();
	uint64_t excl = incl - sum + (warp > 0 ? warp_sums[warp - 1] : 0);
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; i++) {
		if (base + i <= n) out[base + i] = excl;      /* slot n receives the grand total */
		excl += v[i];
	}
	if (threadIdx.x == SCAN_T - 1 && tile_sums != nullptr) tile_sums[blockIdx.x] = excl;
}

__global__ void
scan_add_kernel(uint64_t *out, const uint64_t *tile_offsets, uint64_t n)
{
	const uint64_t i = (uint64_t) blockIdx.x * SCAN_TILE + threadIdx.x;
	const uint64_t add = tile_offsets[blockIdx.x];
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; k++) {
		const uint64_t j = i + (uint64_t) k * SCAN_T;
		if (j < n) out[j] += add;
	}
}

struct Scanner {
	DBuf<uint64_t> lvl[4];
	cudaStream_t st;
	/* out[i] = sum(in[0..i)), out[n] = total (out must hold n+1 entries) */
	template <typename TIn>
	int run(const TIn *in, uint64_t *out, uint64_t n, int depth = 0) {
		if (n == 0) {
			CK(cudaMemsetAsync(out, 0, sizeof(uint64_t), st));
			return 0;
		}
		const uint64_t m = n + 1;                      /* scan n+1 items: the extra slot yields the total */
		const uint64_t tiles = (m + SCAN_TILE - 1) / SCAN_TILE;
		if (lvl[depth].reserve(2 * tiles + 2, false, st) != 0) return -1;
		uint64_t *sums = lvl[depth].p, *sums_scanned = lvl[depth].p + tiles + 1;
		scan_tile_kernel<TIn><<<(unsigned) tiles, SCAN_T, 0, st>>>(in, out, sums, n);   /* items >= n read as 0 */
		count_launch();
		if (tiles > 1) {
			if (depth >= 3) { set_error("determinise: scan too deep"); errno = EIO; return -1; }
			if (run<uint64_t>(sums, sums_scanned, tiles, depth + 1) != 0) return -1;
			scan_add_kernel<<<(unsigned) tiles, SCAN_T, 0, st>>>(out, sums_scanned, m);
			count_launch();
		}
		return 0;
	}
};

__device__ __forceinline__ uint64_t
mix64(uint64_t h, uint64_t v)
{
	h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
	h *= 0xff51afd7ed558ccdull;
	return h ^ (h >> 33);
}

/* class mask of every group: bit k set iff the group's label set contains class k's
 * representative symbol (hence the whole class) */
__global__ void
k2_group_classmask_kernel(const uint64_t *gsym, uint32_t ngroups, const uint8_t *rep, uint32_t K, uint64_t *gcls)
{
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= ngroups) return;
	uint64_t m[4] = { 0, 0, 0, 0 };
	for (uint32_t k = 0; k < K; k++) {
		const uint32_t c = rep[k];
		if ((gsym[4 * (size_t) g + (c >> 6)] >> (c & 63)) & 1ull) m[k >> 6] |= 1ull << (k & 63);
	}
	gcls[4 * (size_t) g + 0] = m[0]; gcls[4 * (size_t) g + 1] = m[1];
	gcls[4 * (size_t) g + 2] = m[2]; gcls[4 * (size_t) g + 3] = m[3];
}

__global__ void
k2_fill_u32_kernel(uint32_t *p, uint32_t v, uint64_t n)
{
	const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = v;
}

__global__ void
k2_fill_u64_kernel(uint64_t *p, uint64_t v, uint64_t n)
{
	const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = v;
}

/* emit: per DFA state, one group per distinct destination, ascending (edge_set keeps groups
 * sorted by .to, src/adt/edgeset.c:283-373); label set = union of the classes' symbols */
__global__ void
k2_emit_count_kernel(const uint32_t *trans, uint32_t D, uint32_t K, uint32_t *ngroups)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= D) return;
	const uint32_t *row = trans + (size_t) s * K;
	uint32_t cnt = 0;
	for (uint32_t k = 0; k < K; k++) {
		const uint32_t to = row[k];
		if (to == NONE32) continue;
		bool seen = false;
		for (uint32_t j = 0; j < k && !seen; j++) seen = row[j] == to;
		cnt += seen ? 0u : 1u;
	}
	ngroups[s] = cnt;
}

__global__ void
k2_emit_fill_kernel(const uint32_t *trans, uint32_t D, uint32_t K, const uint64_t *class_mask,
	const uint64_t *goff, uint32_t *gto, uint64_t *gsym)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= D) return;
	const uint32_t *row = trans + (size_t) s * K;
	uint64_t o = goff[s];
	const uint64_t e = goff[s + 1];
	uint32_t last = 0;
	bool first = true;
	for (; o < e; o++) {
		uint32_t best = NONE32;                       /* next destination in ascending order */
		for (uint32_t k = 0; k < K; k++) {
			const uint32_t to = row[k];
			if (to != NONE32 && (first || to > last) && to < best) best = to;
		}
		uint64_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
		for (uint32_t k = 0; k < K; k++) {
			if (row[k] == best) {
				m0 |= class_mask[4 * k]; m1 |= class_mask[4 * k + 1];
				m2 |= class_mask[4 * k + 2]; m3 |= class_mask[4 * k + 3];
			}
		}
		gto[o] = best;
		gsym[4 * o] = m0; gsym[4 * o + 1] = m1; gsym[4 * o + 2] = m2; gsym[4 * o + 3] = m3;
		last = best;
		first = false;
	}
}

inline unsigned
blocks_for(uint64_t n, unsigned t = 256)
{
	return (unsigned) ((n + t - 1) / t);
}

double
ms_since(std::chrono::steady_clock::time_point t0)
{
	return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

/* Partition 0..255 into classes no label set distinguishes. */
uint32_t
byte_classes(const fsm_b200_desc *d, uint64_t ngroups, uint8_t class_of[256], uint8_t rep[256])
{
	std::map<std::array<uint64_t, 4>, int> distinct;
	for (uint64_t g = 0; g < ngroups; g++) {
		std::array<uint64_t, 4> key = { d->group_symbols[4 * g], d->group_symbols[4 * g + 1],
		                                d->group_symbols[4 * g + 2], d->group_symbols[4 * g + 3] };
		distinct.emplace(key, 0);
	}
	uint64_t sig[256];
	for (int c = 0; c < 256; c++) sig[c] = 0;
	for (auto &kv : distinct) {
		for (int c = 0; c < 256; c++) {
			const uint64_t bit = (kv.first[c >> 6] >> (c & 63)) & 1ull;
			sig[c] = (sig[c] ^ (bit + 0x9e3779b97f4a7c15ull)) * 0xff51afd7ed558ccdull;
			sig[c] ^= sig[c] >> 29;
		}
	}
	/* exact grouping: compare membership vectors, not just hashes */
	std::vector<std::vector<uint8_t>> member(256);
	for (int c = 0; c < 256; c++) {
		member[c].reserve(distinct.size());
		for (auto &kv : distinct) member[c].push_back((uint8_t) ((kv.first[c >> 6] >> (c & 63)) & 1ull));
	}
	uint32_t K = 0;
	for (int c = 0; c < 256; c++) {
		int found = -1;
		for (uint32_t k = 0; k < K; k++) {
			if (sig[rep[k]] == sig[c] && member[rep[k]] == member[c]) { found = (int) k; break; }
		}
		if (found < 0) { rep[K] = (uint8_t) c; found = (int) K; K++; }
		class_of[c] = (uint8_t) found;
	}
	return K;
}




} // namespace

#endif
