/*
 * ref_harness.h -- TEST INFRASTRUCTURE.  Thin helper over the UNMODIFIED reference
 * (oracle/_ref/libfsm_ref.so, compiled from /root/reference by oracle/Makefile).
 * Lets Python tests and bench.py (cpu_baseline / --impl reference) build automata with
 * the reference's own re_comp / fsm_determinise / fsm_minimise / fsm_union_array, dump
 * them into the flat `struct fsm_b200_desc`, and run the reference's own fsm_exec.
 * Handles are `struct fsm *` of the reference passed as void *.
 */
#ifndef REF_HARNESS_H
#define REF_HARNESS_H

#include <stddef.h>
#include <stdint.h>
#include "../include/fsm_b200.h"

/* re_comp(dialect, ...) (include/re/re.h:136-140). dialect: enum re_dialect value
 * (RE_PCRE = 5, RE_NATIVE = 3, RE_LITERAL = 1, RE_GLOB = 2). Returns NULL on error. */
void *refh_re_comp(const char *pattern, size_t len, int dialect, int flags);
void *refh_parse_file(const char *path);                /* fsm_parse of an fsm(5) file */
int   refh_determinise(void *fsm);                       /* fsm_determinise: 1 ok */
int   refh_determinise_limit(void *fsm, size_t state_limit); /* enum ..._res value */
int   refh_minimise(void *fsm);                          /* fsm_minimise: 1 ok */
int   refh_setendid(void *fsm, unsigned id);             /* fsm_setendid */
void *refh_union_array(size_t n, void **fsms);           /* fsm_union_array; consumes inputs */
void *refh_clone(const void *fsm);
void  refh_free(void *fsm);
unsigned refh_countstates(const void *fsm);
int   refh_equal(const void *a, const void *b);          /* fsm_equal */
int   refh_remove_epsilons(void *fsm);

/* Build a reference `struct fsm` from a flat description (fsm_new, fsm_addstate_bulk,
 * fsm_addedge_literal / fsm_addedge_epsilon, fsm_setend, fsm_endid_set, fsm_setstart). */
void *refh_from_desc(const struct fsm_b200_desc *d);

/* Flatten a reference `struct fsm` (arrays malloc'd; release with refh_flat_free). */
struct refh_flat {
	struct fsm_b200_desc desc;
	void *blocks[8];
};
int  refh_flatten(const void *fsm, struct refh_flat *out);
void refh_flat_free(struct refh_flat *f);

/* Epsilon closure through the reference's epsilon_closure() (closure.c:130-190):
 * CSR, malloc'd. */
int refh_epsilon_closure(void *fsm, uint64_t **off, uint32_t **to);

/* One reference fsm_exec call over buf[0..len) through a length-bounded getc callback.
 * out->ret is fsm_exec's return; out->end is *end when ret==1 else UINT32_MAX;
 * out->consumed is derived from the number of getc calls. */
int refh_exec(const void *fsm, const uint8_t *buf, uint64_t len, struct fsm_b200_result *out);

/* n fsm_exec calls over nthreads pthreads.
 * mode 0 "as-is":     the reference's fsm_exec per string (validates DFA-ness each call)
 * mode 1 "amortised": validation hoisted; per byte the reference's own
 *                     edge_set_transition (src/adt/edgeset.c:565-579) + fsm_isend; also
 *                     fills out->end for ret==0 (the state at which the walk stopped). */
int refh_exec_batch(const void *fsm, const uint8_t *base, const uint64_t *offsets, size_t n,
	int mode, int nthreads, struct fsm_b200_result *out);

/* fsm_endid_count / fsm_endid_get for one state; returns count, fills up to cap ids. */
size_t refh_endids(const void *fsm, unsigned state, unsigned *ids, size_t cap);

/* Eager outputs (include/fsm/fsm.h:273-336): set one, dump all as a CSR (sorted unique per
 * state, malloc'd), run fsm_exec with a collecting callback (ids ascending, *nfired may exceed
 * cap), and the reference's own way of attaching them to a union of patterns. */
int   refh_eager_set(void *fsm, unsigned state, unsigned id);
int   refh_eager_flatten(const void *fsm, uint64_t **off, uint32_t **ids);
int   refh_exec_eager(void *fsm, const uint8_t *buf, uint64_t len, struct fsm_b200_result *out,
	unsigned *fired, size_t cap, size_t *nfired);
void *refh_union_repeated_pattern_group(size_t n, void **fsms, unsigned id_base);
/* n fsm_exec calls with the callback installed, nthreads pthreads (each on its own fsm_clone: the
 * callback slot lives in the fsm).  masks[i * words ..): bit b <=> id_of_bit[b] (ascending) fired on
 * line i.  mode 0: the reference's fsm_exec as-is; mode 1: validation hoisted, the reference's own
 * edge_set_transition + fsm_eager_output_iter_state per byte.  out[i].end is filled for ret == 0 too. */
int   refh_exec_eager_batch(void *fsm, const uint8_t *base, const uint64_t *offsets, size_t n,
	int mode, int nthreads, struct fsm_b200_result *out, uint64_t *masks, size_t words,
	const uint32_t *id_of_bit, size_t nbits);
double refh_last_walk_seconds(void);

/* examples/utf8dfa/main.c over the same API calls: the DFA of the code points lo..hi (one code point
 * per input), determinised + minimised; and the Kleene star (epsilon from every end state to the start
 * state, start becomes an end state; the caller determinises + minimises): BASELINE config 4's validator. */
void *refh_utf8dfa(int lo, int hi);
int   refh_star(void *fsm);

/* The reference's bytecode engine: the "DFAVM$" image of a DFA (fsm_vm_compile + fsm_dfavm_save; malloc'd),
 * and n fsm_vm_match_buffer calls over nthreads pthreads (bool only: vm.c:218-229). */
int refh_dfavm_bytes(const void *fsm, uint8_t **out, size_t *len);
int refh_vm_match_batch(const void *fsm, const uint8_t *base, const uint64_t *offsets, size_t n, int nthreads, uint8_t *out);

#endif
