/* k1_device.cuh -- device helpers shared by the K1 kernels: mbarrier / TMA bulk staging of a table
 * blob into shared memory, 256-bit sector loads. */
#ifndef FSM_B200_K1_DEVICE_CUH
#define FSM_B200_K1_DEVICE_CUH

#include <cstdint>

namespace fsmb200 {

__device__ __forceinline__ uint32_t
smem_u32(const void *p)
{
	return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void
mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}

__device__ __forceinline__ void
mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void
mbar_wait(uint32_t bar, uint32_t parity)
{
	uint32_t done;
	do {
		asm volatile(
		    "{\n\t.reg .pred p;\n\t"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		    "selp.u32 %0, 1, 0, p;\n\t}"
		    : "=r"(done) : "r"(bar), "r"(parity) : "memory");
	} while (!done);
}

/* 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP). */
__device__ __forceinline__ void
tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
	asm volatile(
	    "cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	    :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

/* Stage a table blob into shared memory with TMA bulk copies issued by one thread; everybody
 * waits on the mbarrier.  blob_bytes is a multiple of 16. */
__device__ __forceinline__ void
stage_blob(uint8_t *smem, const uint8_t *blob, uint32_t blob_bytes, uint64_t *bar)
{
	const uint32_t bar_a = smem_u32(bar);
	if (threadIdx.x == 0) {
		mbar_init(bar_a, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		mbar_expect_tx(bar_a, blob_bytes);
		const uint32_t dst = smem_u32(smem);
		for (uint32_t off = 0; off < blob_bytes; off += 16384u) {
			const uint32_t nb = min(16384u, blob_bytes - off);
			tma_bulk_g2s(dst + off, blob + off, nb, bar_a);
		}
	}
	mbar_wait(bar_a, 0);
}

/* the same with an L2 prefetch-size hint: a miss makes L2 fetch 256 B (SASS: LDG.E.ENL2.LTC256B.256) */
__device__ __forceinline__ void
ld256_l2_256(const uint8_t *p, uint32_t (&w)[8])
{
	asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	    : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]),
	      "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
	    : "l"(p));
}

/* one full 32-byte DRAM sector per lane (SASS: LDG.E.ENL2.256) */
__device__ __forceinline__ void
ld256(const uint8_t *p, uint32_t (&w)[8])
{
	asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	    : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]),
	      "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
	    : "l"(p));
}

} // namespace fsmb200
#endif
