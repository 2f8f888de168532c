/*
 * fsm_b200_shim.h -- the libfsm-facing side of the drop-in boundary.
 *
 * This file and fsm_b200_shim.c are the pieces a libfsm maintainer adds to the reference
 * tree: they are compiled against the reference's own headers
 * (<fsm/fsm.h>, "libfsm/internal.h", <adt/edgeset.h>, <adt/stateset.h>) and REPLACE
 * src/libfsm/exec.c in libfsm.  re(1), fsm(1), rx(1), lx(1) and every other caller relink
 * unchanged: `fsm_exec` keeps its signature and conventions (include/fsm/fsm.h:560-562).
 */
#ifndef FSM_B200_SHIM_H
#define FSM_B200_SHIM_H

#include <stddef.h>
#include <stdint.h>

#include <fsm/fsm.h>

#include "fsm_b200.h"

/* Flatten a `struct fsm` into the engine's description.  Arrays are malloc'd and owned by
 * `out`; release with fsm_b200_flat_free.  Returns 0, or -1/ENOMEM. */
struct fsm_b200_flat {
	struct fsm_b200_desc desc;      /* desc.reserved has FSM_B200_DESC_EAGER iff the fsm has eager outputs; */
	const uint64_t *eager_off;      /* then &desc is also a struct fsm_b200_desc_ext * (same layout) */
	const uint32_t *eager_ids;
	void *blocks[10];
};
int  fsm_b200_flatten(const struct fsm *fsm, struct fsm_b200_flat *out);
void fsm_b200_flat_free(struct fsm_b200_flat *flat);

/* Additive batch entry: n independent fsm_exec calls in one go.
 * Input i is base[offsets[i] .. offsets[i+1]).  Returns 0, or -1 with errno as fsm_exec
 * (EINVAL: not a DFA / no start state). */
int fsm_exec_batch(const struct fsm *fsm, const unsigned char *base, const uint64_t *offsets,
	size_t n, struct fsm_b200_result *out);

/* Additive: the batch form for automata with eager outputs (fsm_eager_output_set,
 * include/fsm/fsm.h:273-336; produced by fsm_union_repeated_pattern_group).  As fsm_exec_batch,
 * plus per input the bitset of eager-output ids fired along its walk: masks[i*words + b/64] bit
 * b%64 <=> id_of_bit[b] fired, words = (*nbits + 63) / 64 (the caller sizes masks for
 * FSM_B200_EAGER_MAX_IDS / 64 words per input, or calls once with n == 0 to learn *nbits).
 * id_of_bit is CALLER storage for FSM_B200_EAGER_MAX_IDS entries; the first *nbits are filled with
 * the distinct eager-output ids of the automaton, ascending (a copy: nothing the library owns is
 * handed out, so the table cache may evict the compiled automaton at any time). */
int fsm_exec_batch_eager(const struct fsm *fsm, const unsigned char *base, const uint64_t *offsets,
	size_t n, struct fsm_b200_result *out, uint64_t *masks, uint32_t *nbits, uint32_t *id_of_bit);

/* Drop the cached device table of `fsm` (call from fsm_free and from mutators; the shim
 * also revalidates a cheap fingerprint on every call, so this is an optimisation). */
void fsm_b200_invalidate(const struct fsm *fsm);

#endif
