/* k1_exec_batch.h -- internal interface of the K1 batch kernels. */
#ifndef FSM_B200_K1_H
#define FSM_B200_K1_H

#include "common.h"

namespace fsmb200 {

enum K1Variant {
	K1_AUTO = 0,
	K1_LANE = 1,       /* lane-per-input, 256-bit global loads */
	K1_TILE64 = 2,     /* warp tile 32 x 64 B, 2-stage TMA ring */
	K1_TILE32 = 3,     /* warp tile 32 x 32 B, 4-stage TMA ring */
	K1_TILE128 = 4,    /* warp tile 32 x 128 B, 2-stage TMA ring (fewer warps) */
	K1_TILE64x3 = 5,   /* warp tile 32 x 64 B, 3-stage TMA ring (fewer warps) */
	K1_KSTRIDE = 6,    /* lane-per-input, one table lookup per K bytes via byte-class tuples */
	K1_VARIANT_COUNT
};

struct K1Args {
	const uint8_t *base;
	const uint64_t *offsets;    /* n+1 entries, or nullptr for fixed stride */
	const uint64_t *ends;       /* optional: input i = [offsets[i], ends[i]) (K1b jobs) */
	const uint32_t *entry;      /* optional: per-input entry state (K1b jobs), else `start` */
	uint64_t stride, len;
	uint64_t n;
	const uint32_t *n_dev;      /* optional: actual input count lives on the device (<= n) */
	fsm_b200_result *out;        /* result i -> out[i] ... */
	fsm_b200_result *peer_out[7]; /* ... and, fused gather, -> peer_out[r][i] over NVLink P2P */
	uint32_t npeers;
	uint32_t peer_compact;
	/* completion signal of the fused gather: the last CTA to finish (after a system-scope fence
	 * in every CTA) stores sig_value into one flag word in every peer's memory */
	uint32_t *sig_counter;        /* local, zero between launches */
	uint32_t *sig_flags[8];       /* [npeers + 1]: peers' flag words, then this rank's own */
	uint32_t sig_value;        /* 1: peers receive 4-byte match ids (ret << 31 | end) instead of 16-byte records */
	const uint8_t *blob;        /* table rows then is_end bytes */
	uint32_t blob_bytes;
	uint32_t is_end_off;
	uint32_t cls_off;           /* class LUT (CLS kernels) */
	uint32_t pitch;
	uint32_t start;
	uint32_t dead;
	/* k-stride kernels */
	const uint8_t *kblob;
	uint32_t kblob_bytes, kpitch, k1pitch, k1_off, kend_off, klut_off;
	uint32_t kr_add_lo[2], kr_add_hi[2], kr_hxor[2];   /* k-range kernel: the two cell ranges (dfa_compile.cu) */
	uint32_t kr_prefetch;       /* k-range kernel: L2 prefetch distance in bytes (0: none) */
	const uint8_t *absorb;      /* [ntable] 1 = all 256 edges loop back (ragged kernel: stop reading the line) */
	uint32_t prefer_lane;       /* K1b jobs: long inputs, keep the 3-instructions-per-byte LANE kernel */
	uint32_t tile_stage_off;    /* TILE variants: shared-memory carve-up */
	uint32_t tile_bar_off;
	uint32_t tile_main_warps;  /* k-range tile kernel: warps per CTA that run the main rounds (0: all) */
	uint32_t tile_main_rounds; /* ... and how many rounds they run before the all-warp final rounds */
};

bool k1_tile_eligible(const fsm_b200_dfa *dfa, const uint8_t *d_base, const uint64_t *d_offsets,
	uint64_t stride, uint64_t len, size_t n);

int k1_launch(const fsm_b200_dfa *dfa, const uint8_t *d_base, const uint64_t *d_offsets,
	uint64_t stride, uint64_t len, size_t n, fsm_b200_result *d_out, cudaStream_t stream, int variant,
	fsm_b200_result *const *peer_outs = nullptr, int npeers = 0, int peer_compact = 0,
	uint32_t *sig_counter = nullptr, uint32_t *const *sig_flags = nullptr, uint32_t sig_value = 0);

/* K1b jobs: input i = d_base[d_begs[i] .. d_ends[i]) walked from state d_entry[i] (LANE variant). */
int k1_launch_jobs(const fsm_b200_dfa *dfa, const uint8_t *d_base, const uint64_t *d_begs,
	const uint64_t *d_ends, const uint32_t *d_entry, size_t n_max, const uint32_t *d_n,
	fsm_b200_result *d_out, cudaStream_t stream);

/* Ragged / eager batches on shared-memory-resident tables (k1_lines.cu). */
bool k1_lines_eligible(const fsm_b200_dfa *dfa);
int k1_lines_launch(const fsm_b200_dfa *dfa, const uint8_t *d_base, const uint64_t *d_offsets,
	uint64_t stride, uint64_t len, size_t n, fsm_b200_result *d_out, uint64_t *d_masks, cudaStream_t stream,
	const uint32_t *d_entry = nullptr);
/* k1b_stream.cu: one long input on an automaton with eager outputs (records + the set of fired ids) */
bool k1b_stream_eager_ok(const fsm_b200_dfa *dfa, uint64_t len);
int k1b_exec_stream_eager(const fsm_b200_dfa *dfa, const uint8_t *d_buf, uint64_t len,
	fsm_b200_result *h_rec, uint64_t *h_masks, cudaStream_t stream);

} // namespace fsmb200
#endif
