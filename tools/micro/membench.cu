// membench.cu -- what read bandwidth does the K1 access pattern allow?  (tools/micro: measurement aid,
// not part of the library.)  Every lane owns a 1 KiB row and reads it front to back, G bytes per
// step (G = 32: one 256-bit load; 64 / 128: 2 / 4 back-to-back loads of the same 128-byte line), with
// D steps in flight; "coalesced" = a warp reads 1 KiB contiguous per step.  Prints GB/s per variant.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld256(const uint8_t *p, uint32_t (&w)[8]) {
	asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	    : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}

template <int G, int D>       // G bytes per lane per step (32/64/128), D steps in flight
__global__ void __launch_bounds__(1024, 1) rows_kernel(const uint8_t *base, uint64_t nrows, uint32_t *out) {
	constexpr int V = G / 32;
	uint32_t acc = 0;
	const uint64_t nthreads = (uint64_t) gridDim.x * blockDim.x;
	for (uint64_t r = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += nthreads) {
		const uint8_t *p = base + r * 1024;
		uint32_t buf[D][V][8];
#pragma unroll
		for (int d = 0; d < D - 1; d++)
#pragma unroll
			for (int v = 0; v < V; v++) ld256(p + (d * V + v) * 32, buf[d][v]);
		constexpr int STEPS = 1024 / G;
#pragma unroll
		for (int s = 0; s < STEPS; s++) {
			const int slot = s % D, ahead = s + D - 1;
			if (ahead < STEPS) {
#pragma unroll
				for (int v = 0; v < V; v++) ld256(p + (ahead * V + v) * 32, buf[ahead % D][v]);
			}
#pragma unroll
			for (int v = 0; v < V; v++)
#pragma unroll
				for (int k = 0; k < 8; k++) acc += buf[slot][v][k];
		}
	}
	if (acc == 0x12345678u) out[0] = acc;
}

// groups of GL lanes read GL*32 contiguous bytes of one row per instruction; a lane group owns GL rows
// and visits them round-robin (line by line), D instructions in flight: the access pattern of a
// "GL lanes transpose GL rows" scheme
template <int GL, int D>
__global__ void __launch_bounds__(1024, 1) group_kernel(const uint8_t *base, uint64_t nrows, uint32_t *out) {
	uint32_t acc = 0;
	const uint64_t nthreads = (uint64_t) gridDim.x * blockDim.x;
	const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t sub = (uint32_t) (tid % GL);
	for (uint64_t g0 = (tid / GL) * GL; g0 < nrows; g0 += nthreads) {      // this group's GL rows: g0 .. g0+GL-1
		constexpr int CH = 1024 / (GL * 32);                                // chunks of GL*32 B per row
		constexpr int STEPS = CH * GL;                                      // (chunk, row) pairs
		uint32_t buf[D][8];
#pragma unroll
		for (int d = 0; d < D - 1; d++) ld256(base + (g0 + d % GL) * 1024 + (uint64_t) (d / GL) * GL * 32 + sub * 32, buf[d]);
#pragma unroll 8
		for (int s = 0; s < STEPS; s++) {
			const int ahead = s + D - 1;
			if (ahead < STEPS) ld256(base + (g0 + ahead % GL) * 1024 + (uint64_t) (ahead / GL) * GL * 32 + sub * 32, buf[ahead % D]);
#pragma unroll
			for (int k = 0; k < 8; k++) acc += buf[s % D][k];
		}
	}
	if (acc == 0x12345678u) out[0] = acc;
}

__global__ void __launch_bounds__(1024, 1) coalesced_kernel(const uint4 *base, uint64_t n16, uint32_t *out) {
	uint32_t acc = 0;
	const uint64_t nthreads = (uint64_t) gridDim.x * blockDim.x;
	uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	for (; i + 3 * nthreads < n16; i += 4 * nthreads) {
		const uint4 a = __ldg(base + i), b = __ldg(base + i + nthreads), c = __ldg(base + i + 2 * nthreads), d = __ldg(base + i + 3 * nthreads);
		acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
	}
	if (acc == 0x12345678u) out[0] = acc;
}

template <typename F> static float timeit(F f) {
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int i = 0; i < 3; i++) f();
	cudaEventRecord(e0);
	for (int i = 0; i < 10; i++) f();
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	return ms / 10;
}

int main() {
	const uint64_t nrows = 1ull << 20, bytes = nrows * 1024;
	uint8_t *d; uint32_t *o;
	cudaMalloc(&d, bytes); cudaMalloc(&o, 64); cudaMemset(d, 1, bytes);
	int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
#define RUN(G, D) { float ms = timeit([&] { rows_kernel<G, D><<<sms, 1024>>>(d, nrows, o); }); \
	printf("{\"pattern\": \"lane-per-row\", \"bytes_per_step\": %d, \"steps_in_flight\": %d, \"GBps\": %.1f}\n", G, D, bytes / ms / 1e6); }
	RUN(32, 1) RUN(32, 2) RUN(32, 3) RUN(32, 4) RUN(64, 1) RUN(64, 2) RUN(64, 3) RUN(128, 1) RUN(128, 2)
#define RUNG(GL, D) { float ms = timeit([&] { group_kernel<GL, D><<<sms, 1024>>>(d, nrows, o); }); \
	printf("{\"pattern\": \"%d lanes read %d contiguous bytes of a row, %d rows round-robin\", \"in_flight\": %d, \"GBps\": %.1f}\n", GL, GL * 32, GL, D, bytes / ms / 1e6); }
	RUNG(2, 2) RUNG(2, 4) RUNG(4, 2) RUNG(4, 4) RUNG(8, 2) RUNG(8, 4) RUNG(16, 4) RUNG(32, 4)
	{ float ms = timeit([&] { coalesced_kernel<<<sms, 1024>>>((const uint4 *) d, bytes / 16, o); });
	  printf("{\"pattern\": \"coalesced 16 B per lane, 4 in flight\", \"GBps\": %.1f}\n", bytes / ms / 1e6); }
	{ float ms = timeit([&] { coalesced_kernel<<<sms * 2, 1024>>>((const uint4 *) d, bytes / 16, o); });
	  printf("{\"pattern\": \"coalesced 16 B per lane, 4 in flight, 2 CTAs per SM\", \"GBps\": %.1f}\n", bytes / ms / 1e6); }
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
	return 0;
}
