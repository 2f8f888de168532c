/*
 * fsm_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference hot path (katef/libfsm fsm_exec and
 * fsm_determinise) over the flat `struct fsm_b200_desc`.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this; libfsm_b200.so never links or calls it.
 *
 * Parity is PINNED: tests/test_oracle_reference.py checks every function here against
 * the unmodified reference compiled into oracle/_ref/libfsm_ref.so (same DFAs, same
 * inputs, bit-exact records) and against the golden fixtures in tests/golden/ that were
 * generated from the reference by tests/golden/make_golden.py.
 */
#ifndef FSM_ORACLE_H
#define FSM_ORACLE_H

#include "../include/fsm_b200.h"   /* struct fsm_b200_desc / fsm_b200_result only */

/* fsm_all(fsm, fsm_isdfa) && fsm_getstart: 1 if DFA with a start state, else 0.
 * (src/libfsm/walk/all.c:15-31, pred/isdfa.c:25-55, src/adt/edgeset.c:514-562,
 *  src/libfsm/start.c:34-48) */
int oracle_isdfa(const struct fsm_b200_desc *d);

/* One fsm_exec call (src/libfsm/exec.c:85-167) over buf[0..len).
 * Returns ret (1/0, or -1 with errno=EINVAL when not a DFA); fills *out as
 * include/fsm_b200.h documents.  validate=0 skips the per-call fsm_isdfa walk (the
 * "amortised" baseline of BASELINE.md section 3). */
int oracle_exec(const struct fsm_b200_desc *d, const uint8_t *buf, uint64_t len,
	int validate, struct fsm_b200_result *out);

/* fsm_exec with eager outputs (exec.c:55-83,126-130,140-144): the SET of ids fired along the
 * walk -- those of the start state, then of every state entered -- ascending, whether or not the
 * input matches.  `d` may be a plain desc (no ids) or a struct fsm_b200_desc_ext.  *nfired may
 * exceed cap (only cap ids are stored).  Returns oracle_exec's value (validate = 1). */
int oracle_exec_eager(const struct fsm_b200_desc *d, const uint8_t *buf, uint64_t len,
	struct fsm_b200_result *out, uint32_t *fired, size_t cap, size_t *nfired);

/* n independent fsm_exec calls, strings partitioned over nthreads pthreads.
 * Returns 0, or -1/EINVAL if not a DFA. validate_each: as oracle_exec's validate. */
int oracle_exec_batch(const struct fsm_b200_desc *d, const uint8_t *base,
	const uint64_t *offsets, size_t n, int validate_each, int nthreads,
	struct fsm_b200_result *out);

/* Dense table [nstates][256], UINT32_MAX = no edge; first matching group wins
 * (edge_set_find, src/adt/edgeset.c:394-418). */
void oracle_flatten(const struct fsm_b200_desc *d, uint32_t *table);

/* Epsilon closure of every state (closure.c:130-190): CSR out, closures sorted, each
 * including the state itself.  Caller frees *off and *to with free(). */
int oracle_epsilon_closure(const struct fsm_b200_desc *d, uint64_t **off, uint32_t **to);

/* Subset construction (determinise.c:23-335 after epsilons.c:121-288), textbook
 * formulation: DFA state 0 = closure(start); states numbered in BFS discovery order over
 * symbols 0..255.  Result arrays are malloc'd; free with oracle_desc_free.
 * Eager outputs (input given as struct fsm_b200_desc_ext): a DFA state carries the union of the
 * ids of every state in the epsilon closure of every member (epsilons.c:221-253 then
 * determinise.c:2614-2636).
 * Returns 0 ok, 1 state limit reached, -1 errno. */
struct oracle_owned_desc {
	struct fsm_b200_desc desc;
	void *blocks[8];
	/* eager-output sets of the result (NULL when the input carried none): CSR, sorted unique
	 * per state; owned through blocks[6], blocks[7].  desc.reserved stays 0. */
	const uint64_t *eager_off;
	const uint32_t *eager_ids;
};
int oracle_determinise(const struct fsm_b200_desc *nfa, size_t state_limit,
	struct oracle_owned_desc *out);
void oracle_desc_free(struct oracle_owned_desc *d);

/* fsm_minimise (src/libfsm/minimise.c:74-190): trim states that are unreachable from the start
 * or cannot reach an end state (fsm_trim FSM_TRIM_START_AND_END_REACHABLE, minimise.c:93-96),
 * then merge states that cannot be distinguished -- by end-ness, by end-id set
 * (split_ecs_by_end_metadata, minimise.c:733-) or by any label -- with plain Moore partition
 * refinement.  The minimal DFA is unique up to numbering; states are numbered here by the
 * smallest original state of each class.  Input must be a DFA.  Returns 0, or -1 errno. */
int oracle_minimise(const struct fsm_b200_desc *dfa, struct oracle_owned_desc *out);
/* Test hook: the same refinement started from a given initial partition (one class id per input
 * state, 0 = plain non-end state). */
int oracle_minimise_from_classes(const struct fsm_b200_desc *dfa, const uint32_t *cls0, struct oracle_owned_desc *out);

/* Canonical form of a DFA for isomorphism checks: BFS renumbering from the start state
 * following symbols 0..255 in order (unreachable states dropped).
 * canon_table [ncanon][256] (UINT32_MAX = no edge), canon_of_state[nstates]
 * (UINT32_MAX = unreachable), returns ncanon or (uint32_t)-1 if not a DFA.
 * canon_table must have room for nstates*256 entries. */
uint32_t oracle_canonicalise(const struct fsm_b200_desc *d, uint32_t *canon_table,
	uint32_t *canon_of_state);

#endif
