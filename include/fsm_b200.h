/*
 * fsm_b200.h -- C ABI of libfsm_b200.so, the B200 (sm_100a) engine behind libfsm's
 * fsm_exec / fsm_determinise hot path.
 *
 * Everything here is `extern "C"`, plain pointers and sizes: no torch types, no C++.
 * The reference (katef/libfsm) has no FFI or plugin layer for execution -- its boundary
 * is the exported C function
 *
 *     int fsm_exec(const struct fsm *, int (*fsm_getc)(void *), void *opaque,
 *                  fsm_state_t *end, struct fsm_capture *captures);
 *                                                  (reference include/fsm/fsm.h:560-562)
 *
 * plus the in-place transforms fsm_determinise / fsm_determinise_with_config
 * (include/fsm/fsm.h:472-488).  `struct fsm` is an array of per-state heap blocks
 * (src/libfsm/internal.h:47-85) which a GPU cannot consume, so the ABI is split in two:
 *
 *   1. struct fsm_b200_desc: a flat, pointer-to-array description of a `struct fsm`
 *      (states, 256-bit-label edge groups exactly as src/adt/edgeset.c:34-41 stores them,
 *      epsilon sets, end bits, end-id sets).  The libfsm-side shim
 *      (libfsm_b200/shim/fsm_b200_shim.c, built inside the reference tree; see
 *      INTEGRATION.md) produces it by walking edge_set_group_iter
 *      (src/adt/edgeset.c:1226-1338) and calls the functions below; re(1)/fsm(1) relink
 *      against the shim's `fsm_exec` unchanged.
 *
 *   2. the engine entry points below, which only see the flat description.
 *
 * Error convention mirrors the reference: functions returning int give -1 and set errno
 * (EINVAL: not a DFA / no start state, as src/libfsm/exec.c:106-114; ENOMEM; EIO: CUDA
 * failure; ENOTSUP: FSM needs a feature the engine does not accelerate).  There is NO CPU
 * fallback: without a usable sm_100 device every compute entry point fails with EIO.
 */
#ifndef FSM_B200_H
#define FSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: struct fsm_b200_det_stats grew (ms_numbering); struct fsm_b200_desc_ext and the eager-output,
 *    determinise_ex and dfa_plan entry points were added.  Everything of version 1 is unchanged. */
/* 3: struct fsm_b200_dfa_info grew (kclasses, krange*); fsm_exec_batch_eager of the shim copies the
 *    id list into caller storage.  Everything else of version 2 is unchanged. */
/* 4: fsm_b200_exec_stream_map_dev_async and struct fsm_b200_stream_map_entry were added.  Everything of
 *    version 3 is unchanged. */
#define FSM_B200_ABI_VERSION 4

/* ------------------------------------------------------------------------------------
 * Flat description of a `struct fsm` (reference src/libfsm/internal.h:52-85).
 *
 * State s owns edge groups [group_off[s], group_off[s+1]); group g carries a 256-bit
 * label set group_symbols[4*g .. 4*g+3] (bit c of word c/64 == symbol c, the layout of
 * `struct edge_group.symbols`, src/adt/edgeset.c:36-39) and one destination group_to[g].
 * Groups are kept in the reference's stored order (sorted by destination), so "first
 * group whose mask has the symbol wins" (edge_set_find, src/adt/edgeset.c:394-418) is
 * reproducible.  eps_* is the CSR of `struct fsm_state.epsilons`; endid_* the CSR of the
 * per-end-state sorted unique id sets (fsm_endid_get, src/libfsm/endids.c:686-755).
 * eps_off / endid_off may be NULL (== no epsilons / no end ids).
 * ------------------------------------------------------------------------------------ */
struct fsm_b200_desc {
	uint32_t nstates;            /* fsm->statecount */
	uint32_t start;              /* fsm->start, meaningful iff hasstart */
	uint32_t hasstart;           /* fsm->hasstart */
	uint32_t reserved;           /* flags: 0, or FSM_B200_DESC_EAGER (see struct fsm_b200_desc_ext) */
	const uint8_t  *is_end;      /* [nstates] fsm->states[s].end */
	const uint64_t *group_off;   /* [nstates+1] */
	const uint64_t *group_symbols; /* [4*ngroups] */
	const uint32_t *group_to;    /* [ngroups] */
	const uint64_t *eps_off;     /* [nstates+1] or NULL */
	const uint32_t *eps_to;      /* [neps] */
	const uint64_t *endid_off;   /* [nstates+1] or NULL */
	const uint32_t *endids;      /* [nendids], sorted unique per state */
};

/* Extended description: a desc whose `reserved` field has FSM_B200_DESC_EAGER set is the first
 * member of this struct, which adds the CSR of the per-state EAGER OUTPUT id sets
 * (fsm_eager_output_set / fsm_eager_output_get, include/fsm/fsm.h:273-336,
 * src/libfsm/eager_output.c): ids a state emits every time it is entered during fsm_exec
 * (src/libfsm/exec.c:55-83,126-130,140-144), whether or not the input ends up matching.
 * Unlike end ids they may sit on non-end states.  Sets are sorted and unique per state.
 * Old callers (reserved == 0) are unaffected: every entry point reads the extra members only
 * when the flag is set. */
#define FSM_B200_DESC_EAGER 1u
struct fsm_b200_desc_ext {
	struct fsm_b200_desc base;   /* base.reserved & FSM_B200_DESC_EAGER */
	const uint64_t *eager_off;   /* [nstates+1] */
	const uint32_t *eager_ids;   /* [neager], sorted unique per state */
};

/* ------------------------------------------------------------------------------------
 * Per-input result record: everything observable from one reference fsm_exec call.
 *   ret       1 match / 0 no match                         (src/libfsm/exec.c:133-166)
 *   end       ret==1: the accepting state index (`*end`, exec.c:165).
 *             ret==0: the value of the reference's local `state` at return -- the final
 *             non-end state, or the state that had no edge for the next byte.
 *   consumed  value of exec.c's local `offset` at return (exec.c:92,150): the input
 *             length when all input was consumed, else the index of the first byte with
 *             no outgoing edge (the reference stops reading there, exec.c:133-138).
 * The end-id set of a match is endids[endid_off[end] .. endid_off[end+1]) of the desc.
 * ------------------------------------------------------------------------------------ */
struct fsm_b200_result {
	int32_t  ret;
	uint32_t end;
	uint64_t consumed;
};

/* Compiled, device-resident DFA (dense [state][256] transition table in HBM, staged to
 * shared memory by the kernels).  Replaces the per-call fsm_all(fsm, fsm_isdfa)
 * validation of exec.c:106: validation happens once, here. */
typedef struct fsm_b200_dfa fsm_b200_dfa;

/* Library/ABI version (FSM_B200_ABI_VERSION of the built library). */
int fsm_b200_abi_version(void);

/* Number of usable sm_100 devices (0 if none / no driver).  Never fails. */
int fsm_b200_device_count(void);

/* Human-readable description of the last error on this thread (never NULL). */
const char *fsm_b200_last_error(void);

/* Validate `desc` as a DFA exactly as fsm_all(fsm, fsm_isdfa) + fsm_getstart do
 * (src/libfsm/exec.c:106-114, pred/isdfa.c:25-55, src/adt/edgeset.c:514-562) and build
 * the dense table on `device`.  Returns 0 and *out on success; -1/EINVAL if `desc` is not
 * a DFA or has no start state; -1/ENOMEM, -1/EIO otherwise. */
int fsm_b200_dfa_compile(const struct fsm_b200_desc *desc, int device, fsm_b200_dfa **out);

void fsm_b200_dfa_free(fsm_b200_dfa *dfa);

/* Introspection of a compiled DFA. */
struct fsm_b200_dfa_info {
	uint32_t nstates;        /* states of the source fsm */
	uint32_t ntable_states;  /* rows in the dense table (nstates, +1 if a dead row was added) */
	uint32_t start;
	uint32_t entry_bytes;    /* 1, 2 or 4 */
	uint32_t row_pitch_bytes;/* bytes between consecutive rows of the device table */
	uint32_t complete;       /* 1 if every state has all 256 edges (no dead row) */
	uint32_t smem_resident;  /* 1 if the table is staged to shared memory by the kernels */
	uint32_t device;
	uint64_t table_bytes;
	uint32_t nclasses;       /* 0: rows indexed by byte; else by byte class (compressed rows) */
	uint32_t kstride;        /* 0, or K in {2,4}: a K-bytes-per-lookup table is also resident */
	uint32_t kclasses;       /* columns per byte of that table (byte classes, or 4 cell codes when krange != 0) */
	uint32_t krange;         /* 0: the k-stride kernel classifies bytes through shared-memory LUTs; 1 / 2: with
	                          * integer arithmetic in registers -- the byte class is a function of
	                          * [b in R0] + 2 [b in R1] for the two byte ranges below (1: both below 0x80) */
	uint8_t  krange_lo[2], krange_hi[2];   /* R0, R1 (lo > hi: unused) */
	uint32_t lines_smem;     /* 1: ragged batches and eager-output batches run the shared-memory lines kernel */
	uint32_t lines_blob_bytes;/* its blob: 512-byte LUT + class rows (+1 NOP column) + end bytes */
	uint32_t lines_cols;     /* columns per row there (byte classes + 1) */
	uint32_t eager_ids;      /* distinct eager-output ids (0: none) */
};
int fsm_b200_dfa_info(const fsm_b200_dfa *dfa, struct fsm_b200_dfa_info *info);

/* The table layout fsm_b200_dfa_compile WOULD choose for `desc` (entry width, row pitch,
 * shared-memory residency, byte classes, k-stride), computed on the host without touching
 * any device; same validation and errno as fsm_b200_dfa_compile.  info->device = UINT32_MAX. */
int fsm_b200_dfa_plan(const struct fsm_b200_desc *desc, struct fsm_b200_dfa_info *info);

/* Copy the dense table back as uint32 next-state indices, [nstates][256], with
 * UINT32_MAX for "no edge".  For tests of the flattener; not a hot path. */
int fsm_b200_dfa_table(const fsm_b200_dfa *dfa, uint32_t *out /* [nstates*256] */);

/* --- batched execution: n independent inputs == n reference fsm_exec calls ------------
 * Input i is bytes base[offsets[i] .. offsets[i+1]) (any byte values, including 0).
 * `offsets` has n+1 entries, non-decreasing.  out[i] is written for every i.
 *
 * _host: base/offsets/out are HOST pointers; the call copies inputs to the device, runs
 * the kernel and copies the records back (this is the end-to-end path: what a relinked
 * re(1)/fsm(1) reaches through the shim's fsm_exec).
 *
 * _dev: all pointers are DEVICE pointers on the DFA's device; the kernel is enqueued on
 * `stream` (a cudaStream_t passed as void *, NULL = legacy default stream) and the call
 * returns without synchronising.  If d_offsets is NULL the inputs are fixed-stride:
 * input i = d_base[i*stride .. i*stride+len).
 */
int fsm_b200_exec_batch_host(const fsm_b200_dfa *dfa,
	const uint8_t *base, const uint64_t *offsets, size_t n,
	struct fsm_b200_result *out);

int fsm_b200_exec_batch_dev(const fsm_b200_dfa *dfa,
	const uint8_t *d_base, const uint64_t *d_offsets, uint64_t stride, uint64_t len,
	size_t n, struct fsm_b200_result *d_out, void *stream);

/* --- eager outputs ---------------------------------------------------------------------------
 * A DFA compiled from a struct fsm_b200_desc_ext carries per-state eager-output id sets.  The
 * distinct ids of the whole DFA are numbered densely in ascending order ("bits"); one input's
 * answer is the bitset of ids fired along its walk: those of the start state, then of every
 * state entered (exec.c:126-130,140-144), whether or not the input matches -- exactly the set of
 * ids the reference hands to the fsm_eager_output_cb callback (their order and multiplicity are
 * not reproduced; the reference's own tests only use the set, tests/eager_output/utils.c:10-24).
 * fsm_b200_dfa_eager_info: *nbits = number of distinct ids (0: none), *id_of_bit = library-owned
 * array [nbits] valid until fsm_b200_dfa_free.  The mask of input i is masks[i*words ..
 * (i+1)*words), words = (nbits + 63) / 64, bit b of word b/64 set <=> id_of_bit[b] fired.
 * The plain entry points above work on such a DFA too and simply do not report the ids.
 * At most FSM_B200_EAGER_MAX_IDS distinct ids are supported (-1/ENOTSUP at compile time). */
#define FSM_B200_EAGER_MAX_IDS 256u
int fsm_b200_dfa_eager_info(const fsm_b200_dfa *dfa, uint32_t *nbits, const uint32_t **id_of_bit);
int fsm_b200_exec_batch_eager_host(const fsm_b200_dfa *dfa,
	const uint8_t *base, const uint64_t *offsets, size_t n,
	struct fsm_b200_result *out, uint64_t *masks);
int fsm_b200_exec_batch_eager_dev(const fsm_b200_dfa *dfa,
	const uint8_t *d_base, const uint64_t *d_offsets, uint64_t stride, uint64_t len,
	size_t n, struct fsm_b200_result *d_out, uint64_t *d_masks, void *stream);

/* --- multi-GPU: scan fused with the result gather over NVLink peer memory -------------------
 * As _dev, but every record is ALSO stored into npeers (<= 7) peer buffers: peer_outs[r]
 * points at the slot of THIS rank's range inside rank r's gathered buffer (device memory
 * of another GPU of the node, mapped with fsm_b200_ipc_open).  The stores are issued by the
 * scanning lanes themselves (P2P over NVLink/NVSwitch), so the gather overlaps the scan
 * tile by tile and no collective kernel competes for SMs.  After the kernels of all ranks
 * have completed (any cross-rank barrier), every gathered buffer holds every rank's records.
 * compact = 0: peers receive the full 16-byte records (peer_outs[r] is a record array);
 * compact = 1: peers receive 4-byte match ids, (ret == 1) << 31 | end (peer_outs[r] is a
 * uint32_t array) -- a quarter of the NVLink volume; `consumed` stays on the owning rank.
 * Completion signal (optional, sig_counter != NULL): sig_counter is a zeroed uint32 in local
 * device memory; sig_flags[0..npeers-1] point at one uint32 flag word in each peer's memory,
 * sig_flags[npeers] at this rank's own.  When the last CTA of the kernel has finished (all
 * CTAs' peer stores fenced at system scope) it stores sig_value into every flag: a consumer
 * that reads flag == sig_value knows this rank's records of that step are complete -- the
 * handshake costs no collective kernel (which could not co-reside with the persistent scan).
 * Helper entry points: plain cudaMalloc'd buffers (IPC needs whole allocations), handle
 * export/open, and a synchronous read-back for checks. */
int fsm_b200_exec_batch_dev_gather(const fsm_b200_dfa *dfa,
	const uint8_t *d_base, const uint64_t *d_offsets, uint64_t stride, uint64_t len, size_t n,
	struct fsm_b200_result *d_out, void *const *peer_outs, int npeers, int compact,
	void *sig_counter, void *const *sig_flags, uint32_t sig_value, void *stream);
/* The consumer side of that signal: enqueue on `stream` a one-warp kernel that polls d_flags[0..n) (n <= 32
 * flag words in this rank's memory, one per source rank) until each has reached `value` (step numbers
 * only grow), so that whatever is enqueued behind it reads complete records.  Bounded (about a second):
 * on expiry it sets *d_timed_out (optional uint32 in device memory) instead of hanging the device. */
int fsm_b200_wait_flags_dev(int device, const void *d_flags, uint32_t n, uint32_t value, void *d_timed_out, void *stream);
int fsm_b200_dev_alloc(int device, size_t bytes, void **out);
int fsm_b200_dev_free(int device, void *p);
int fsm_b200_dev_zero(int device, void *p, size_t bytes);
int fsm_b200_dev_read(int device, void *host_dst, const void *dev_src, size_t bytes);
int fsm_b200_ipc_export(const void *dev_ptr, void *handle64 /* 64 bytes */);
int fsm_b200_ipc_open(int device, const void *handle64, void **out);
int fsm_b200_ipc_close(int device, void *p);

/* Kernel variant selection for _dev/_host (0 = library default).  Exposed so that
 * bench.py / ncu can evidence the choice; see DESIGN.md "K1 variants". */
int fsm_b200_set_exec_variant(int variant);
int fsm_b200_get_exec_variant(void);

/* --- one long input == one reference fsm_exec call over a stream -----------------------
 * The input is cut into chunks; each chunk computes its state->state map (DFA execution
 * is a monoid), maps are composed by an exclusive scan, and the verdict is identical to a
 * serial walk.  `entry_state` is the state the walk starts in (the DFA's start for a
 * whole input; an arbitrary state for a byte-range shard, see exec_stream_map).
 */
int fsm_b200_exec_stream_host(const fsm_b200_dfa *dfa, const uint8_t *buf, uint64_t len,
	struct fsm_b200_result *out);

int fsm_b200_exec_stream_dev(const fsm_b200_dfa *dfa, const uint8_t *d_buf, uint64_t len,
	struct fsm_b200_result *out /* host */, void *stream);

/* Shard form for multi-GPU range sharding: computes, for the byte range d_buf[0..len),
 * the map entry_state -> (exit_state, first_dead_offset) for EVERY entry state.
 * map_state[s]  = state after consuming the shard from s (dead row index ntable_states-1
 *                 when a missing edge was hit, for incomplete DFAs),
 * map_dead[s]   = offset within the shard of the first byte with no edge, or UINT64_MAX.
 * map_dead_state[s] = state that had no edge for that byte (UINT32_MAX if none).
 * Ranks all-gather these [ntable_states] records and compose them in rank order. */
int fsm_b200_exec_stream_map_dev(const fsm_b200_dfa *dfa, const uint8_t *d_buf, uint64_t len,
	uint32_t *map_state /* host [ntable_states] */,
	uint64_t *map_dead  /* host [ntable_states] */,
	uint32_t *map_dead_state /* host [ntable_states] */, void *stream);

/* The same map left ON THE DEVICE and nothing waited for: d_map (device memory, [nstates] records) is
 * written by work queued on `stream`, so a collective queued behind it on that stream -- the all-gather
 * of the ranks' maps (bench.py --config 4) -- needs no host round trip.  `state` is the exit state, or
 * the state that had no edge when `died`; `dead_off` the shard offset of that byte (UINT64_MAX if none).
 * Calls on one DFA must use one stream (they share scratch memory). */
struct fsm_b200_stream_map_entry {
	uint32_t state;
	uint32_t died;
	uint64_t dead_off;
};
int fsm_b200_exec_stream_map_dev_async(const fsm_b200_dfa *dfa, const uint8_t *d_buf, uint64_t len,
	struct fsm_b200_stream_map_entry *d_map /* device [nstates] */, void *stream);

/* --- determinisation: the subset-construction loop of fsm_determinise -----------------
 * (src/libfsm/determinise.c:23-335 incl. epsilon removal, epsilons.c:121-288).
 * Input: any NFA as a desc.  Output: a DFA as a library-owned desc (free with
 * fsm_b200_desc_free).  State 0 of the output is the start state (determinise.c:234).
 * The result is the reference's DFA up to state renumbering (isomorphic; the reference's
 * numbering is an artefact of its LIFO worklist + pairwise analysis order);
 * end bits and end-id sets are carried as determinise.c:236-266 does.
 * state_limit: 0 = unlimited; otherwise returns 1 (and no output) as soon as more DFA
 * states than the limit would be created (FSM_DETERMINISE_WITH_CONFIG_STATE_LIMIT_REACHED,
 * include/fsm/fsm.h:478-488).  Returns 0 on success, -1/errno on error.
 */
struct fsm_b200_owned_desc {
	struct fsm_b200_desc desc;
	void *owner;             /* opaque; frees every array above */
};
int fsm_b200_determinise(const struct fsm_b200_desc *nfa, int device, size_t state_limit,
	struct fsm_b200_owned_desc *out);
void fsm_b200_desc_free(struct fsm_b200_owned_desc *d);

/* Same, with flags.  FSM_B200_DET_REFERENCE_NUMBERING: number the DFA states exactly as the
 * reference does, i.e. in the order its LIFO worklist (determinise.c:118-185) meets them when
 * every state lists its successors in the entry order of the pairwise label-group analysis
 * (determinise.c:898-1054, :1056-1335, :2331-2505) -- the output is then the reference's
 * `struct fsm` state for state, not merely isomorphic to it.  The per-state successor orders
 * are computed on the device (libfsm_b200/csrc/refnum.h explains the closed form), the
 * worklist walk itself is sequential and runs on the host.  fsm_b200_determinise() uses
 * flags 0 unless the environment says FSM_B200_DET_NUMBERING=reference. */
#define FSM_B200_DET_REFERENCE_NUMBERING 1u
int fsm_b200_determinise_ex(const struct fsm_b200_desc *nfa, int device, size_t state_limit,
	unsigned flags, struct fsm_b200_owned_desc *out);

/* Eager outputs through determinise / minimise: give the input as a struct fsm_b200_desc_ext.
 * A DFA state carries the union of the ids of every state in the epsilon closure of every member
 * (epsilons.c:221-253, determinise.c:2614-2636); minimise keeps states with different id sets
 * apart the way the reference does (same_end_metadata, minimise.c:705-731, including its
 * list-order-dependent blind spot, minimise.c:771-782) and gives a merged state the union
 * (consolidate.c:306-315).  The result's sets are read with this accessor (CSR over the output
 * states, sorted unique; both NULL when there are none); the embedded desc keeps reserved == 0. */
int fsm_b200_owned_desc_eager(const struct fsm_b200_owned_desc *d, const uint64_t **eager_off,
	const uint32_t **eager_ids);

/* Timing of the last determinise on this thread, milliseconds, for bench.py. */
struct fsm_b200_det_stats {
	double ms_total, ms_closure, ms_expand, ms_intern, ms_emit;
	uint64_t dfa_states, dfa_groups, rounds, kernel_launches;
	double ms_numbering;     /* reference numbering (0 when not requested) */
};
int fsm_b200_determinise_stats(struct fsm_b200_det_stats *st);

/* --- DFAVM bytecode loader (host code, no device needed) -------------------------------------
 * A DFA saved in the reference's DFAVM format ("DFAVM$" + encoding 0.1 + u32 length + instruction bytes:
 * fsm_dfavm_save, src/libfsm/vm.c:39-71; vm/v1.c) becomes a library-owned description (free with
 * fsm_b200_desc_free) that fsm_b200_dfa_compile takes like any other: state i = the i-th FETCH of the
 * program, the start state where the program begins; a STOP-success becomes an absorbing accepting state, a STOP-fail a
 * missing edge.  The VM answers yes / no only, so the description carries no end ids.
 * Returns 0, or -1 with errno EINVAL (not a DFAVM image / malformed) or ENOTSUP (other encoding). */
int fsm_b200_dfavm_load(const uint8_t *image, size_t nbytes, struct fsm_b200_owned_desc *out);

/* --- minimisation: the partition refinement of fsm_minimise -----------------------------
 * (src/libfsm/minimise.c:74-190: trim to start- and end-reachable states, then merge
 * indistinguishable states; end states with different end-id sets are never merged,
 * minimise.c:733-).  Input: a DFA as a desc (-1/EINVAL otherwise).  Output: the minimal DFA,
 * unique up to state numbering (classes are numbered by their smallest member here), as a
 * library-owned desc; an automaton that can match nothing comes back with 0 states.
 * Returns 0 on success, -1/errno on error. */
int fsm_b200_minimise(const struct fsm_b200_desc *dfa, int device, struct fsm_b200_owned_desc *out);
int fsm_b200_minimise_stats(struct fsm_b200_det_stats *st);

/* Count of kernel launches issued by this library on this thread since the last reset
 * (bench.py's "gpu_launches"). */
uint64_t fsm_b200_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* FSM_B200_H */
