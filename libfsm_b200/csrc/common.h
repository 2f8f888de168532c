/*
 * common.h -- internal helpers shared by the engine's translation units.
 */
#ifndef FSM_B200_COMMON_H
#define FSM_B200_COMMON_H

#include <cerrno>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/fsm_b200.h"

namespace fsmb200 {

/* thread-local error text + launch counter (api.cu) */
void set_error(const char *fmt, ...);
void count_launch(uint64_t n = 1);

#define FSMB_CUDA(expr, fail_stmt)                                                      \
	do {                                                                                \
		cudaError_t e_ = (expr);                                                        \
		if (e_ != cudaSuccess) {                                                        \
			::fsmb200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
			    __FILE__, __LINE__);                                                    \
			errno = (e_ == cudaErrorMemoryAllocation) ? ENOMEM : EIO;                   \
			fail_stmt;                                                                  \
		}                                                                               \
	} while (0)

constexpr uint32_t NO_EDGE = 0xFFFFFFFFu;

/* Rows of a shared-memory table are padded by 4 bytes so that the bank of an entry
 * depends on the state as well as on the input byte ("bank conflicts"). */
constexpr uint32_t SMEM_ROW_PAD = 4;
/* Largest dense table (bytes, padded) the batch kernels stage into shared memory. */
constexpr uint32_t SMEM_TABLE_MAX = 96 * 1024;
/* Largest byte-class-compressed table staged into shared memory (LANE kernel only: it has
 * no input staging buffers, so nearly all of the 227 KB opt-in limit is available). */
constexpr uint32_t SMEM_CLASS_TABLE_MAX = 200 * 1024;
/* Above this many byte classes the indirection is not worth it for an L2-resident table. */
constexpr uint32_t CLASS_GLOBAL_MAX = 192;
/* Largest lines-kernel blob (LUT + class rows + end bytes) staged into shared memory. */
constexpr uint32_t SMEM_LINES_MAX = 224 * 1024;

/* Pinned host blocks, cached between calls (api.cu).  The big outputs of determinise / minimise (config 5: 94 MB
 * of edge groups) used to land in fresh std::vector storage: 36-50 ms of first-touch page faults per call,
 * then a pageable device-to-host copy.  A block from this pool has its pages faulted in and pinned once; a
 * freed description returns its blocks.  At most 8 free blocks / 1 GiB are kept. */
void *host_pool_get(size_t bytes, size_t *cap);      /* nullptr when pinned memory cannot be had (callers fall back to vectors) */
void host_pool_put(void *p, size_t cap);

/* a pool block for the duration of a scope */
struct PoolTemp {
	void *p = nullptr; size_t cap = 0;
	bool get(size_t bytes) { p = host_pool_get(bytes, &cap); return p != nullptr; }
	~PoolTemp() { if (p != nullptr) host_pool_put(p, cap); }
};

/* Arrays behind a library-owned description (struct fsm_b200_owned_desc.owner). */
struct Owner {
	std::vector<uint8_t> is_end;
	std::vector<uint64_t> group_off, group_sym, endid_off;
	std::vector<uint32_t> group_to, endids;
	/* edge groups in pool blocks instead of the two vectors above (large results) */
	uint32_t *pin_gto = nullptr; size_t pin_gto_cap = 0;
	uint64_t *pin_gsym = nullptr; size_t pin_gsym_cap = 0;
	/* where the groups live: reserve `ng` groups, pool blocks first */
	void groups_alloc(size_t ng) {
		if (ng * 36 >= (1u << 20)) {
			pin_gto = static_cast<uint32_t *>(host_pool_get(ng * 4, &pin_gto_cap));
			pin_gsym = static_cast<uint64_t *>(host_pool_get(ng * 32, &pin_gsym_cap));
			if (pin_gto != nullptr && pin_gsym != nullptr) return;
			if (pin_gto != nullptr) host_pool_put(pin_gto, pin_gto_cap);
			if (pin_gsym != nullptr) host_pool_put(pin_gsym, pin_gsym_cap);
			pin_gto = nullptr; pin_gsym = nullptr;
		}
		group_to.resize(ng);
		group_sym.resize(4 * ng);
	}
	uint32_t *gto() { return pin_gto != nullptr ? pin_gto : group_to.data(); }
	uint64_t *gsym() { return pin_gsym != nullptr ? pin_gsym : group_sym.data(); }
	Owner() = default;
	Owner(const Owner &) = delete;
	Owner &operator=(const Owner &) = delete;
	~Owner() {
		if (pin_gto != nullptr) host_pool_put(pin_gto, pin_gto_cap);
		if (pin_gsym != nullptr) host_pool_put(pin_gsym, pin_gsym_cap);
	}
	/* eager-output sets of the result (fsm_b200_owned_desc_eager); empty when there are none */
	std::vector<uint64_t> eager_off;
	std::vector<uint32_t> eager_ids;
};

} // namespace fsmb200

/* The compiled DFA. */
struct fsm_b200_dfa {
	int device;
	uint32_t nstates;        /* source states */
	uint32_t ntable;         /* table rows = nstates (+1 dead row) */
	uint32_t start;
	uint32_t dead;           /* dead row index or NO_EDGE when complete */
	uint32_t entry_bytes;    /* 1/2/4 */
	uint32_t pitch;          /* row pitch in bytes of d_table */
	uint32_t complete;
	uint32_t smem_resident;
	uint32_t nclasses;       /* 0: rows indexed by byte; else rows indexed by byte class */
	uint32_t is_end_off;     /* blob offsets */
	uint32_t cls_off;
	uint64_t table_bytes;    /* ntable * pitch */
	uint64_t blob_bytes;     /* table | is_end[ntable] | class LUT[256], each padded to 16 */
	uint8_t class_of[256];
	/* k-stride form (tables with <= 256 rows and few byte classes): one lookup per K bytes */
	void *d_kblob;           /* stepK rows | step1 rows | is_end | K class LUTs of 256 B */
	uint32_t kstride;        /* 0 (none), 2 or 4 */
	uint32_t kclasses;
	uint32_t kpitch, k1pitch;            /* row pitches (bytes) of stepK / step1 */
	uint32_t k1_off, kend_off, klut_off; /* blob offsets */
	uint32_t kblob_bytes;
	/* ALU byte classification for the k-stride kernel (0: class LUTs; 1: two ranges below 0x80;
	 * 2: some range at or above 0x80): see find_cell_ranges in dfa_compile.cu */
	uint32_t krange;
	uint32_t kr_add_lo[2], kr_add_hi[2], kr_hxor[2];
	uint8_t kr_lo[2], kr_hi[2];
	void *d_rblob;           /* k-range kernel: [256 B cell LUT][stepK transposed 256 x 256][step1 rows][is_end] */
	uint32_t rblob_bytes, r_k1_off, r_kend_off;
	/* device */
	void *d_blob;            /* table rows followed by is_end bytes (u8 per row) */
	uint8_t *d_absorb;       /* [ntable] 1 = every byte loops back to the state itself */
	uint32_t has_absorbing;  /* some real (non-dead) state is absorbing */
	/* host copies for introspection / stream composition */
	uint32_t *h_table32;     /* [nstates*256], NO_EDGE for missing */
	uint8_t *h_is_end;       /* [ntable] (dead row: 0) */
	/* scratch for the _host entry points (grown on demand; guarded by mutex) */
	void *scratch;
	/* scratch of the stream (K1b) entry points */
	void *stream_scratch;
	/* scratch of fsm_b200_exec_batch_eager_host */
	void *eager_scratch;
	/* eager outputs (k1_eager.cu): distinct ids ascending, per-row bit masks [ntable][eager_words] */
	uint32_t eager_nbits, eager_words;
	uint32_t *h_eager_ids;
	uint64_t *d_eager_masks;
	/* lines kernel (k1_lines.cu; ragged batches and eager outputs on shared-memory-resident tables):
	 * rows indexed by byte class plus one NOP column (every state loops to itself: how bytes outside
	 * a line are walked), states renumbered so that those with eager outputs come last, just before
	 * the dead row: "something to report in this sector" is one max() per byte.
	 * blob: [512-byte LUT: byte -> class, 256..511 -> NOP][rows][is_end by new number] */
	void *d_lblob;
	uint32_t lblob_bytes, l_pitch, l_entry_bytes, l_tab_off, l_end_off, l_first_event, l_dead, l_start, l_ncols;
	uint32_t *d_lperm_inv;   /* [ntable] new number -> caller's state number (dead row -> ntable - 1) */
	uint32_t *d_lperm;       /* [ntable] caller's state number -> new number */
	uint8_t *d_labsorb;      /* [ntable] by new number, or nullptr when no real state is absorbing */
	uint64_t *d_lev_masks;   /* [ntable - l_first_event][eager_words]: id masks of the states with outputs, by new number */
	uint64_t l_start_mask[4];/* eager ids of the start state (exec.c:126-130) */
	uint32_t l_planned;      /* the lines blob exists (fits shared memory) */
};

#endif
