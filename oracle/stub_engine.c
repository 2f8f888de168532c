/*
 * stub_engine.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A CPU stand-in for libfsm_b200.so, exporting exactly the engine symbols the libfsm-side
 * shim (libfsm_b200/shim/fsm_b200_shim.c) calls and answering them with the plain-C oracle
 * (oracle/fsm_oracle.c).  libfsm_b200/shim/Makefile links the UNCHANGED shim source against
 * this into build/shim_cpu/, so that `pytest -m "not gpu"` can exercise the shim's HOST logic
 * -- struct fsm flattening, the device-table cache and its locking, getc draining and cursor
 * restoration, the struct fsm rebuild after determinise/minimise, errno conventions -- with the
 * reference's own CLIs and C unit tests, on a machine without a GPU.
 *
 * It is never shipped and never on a product path: the product shim (build/shim/) links the
 * CUDA engine and fails with EIO when no device is usable.  Every call is counted so the
 * tests can also check WHICH engine entry points a libfsm call reaches.
 */
#define _POSIX_C_SOURCE 200809L
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "fsm_oracle.h"

#define LIVE 0x600DF00Du
#define DEAD 0xDEADDEADu

struct fsm_b200_dfa {
	volatile unsigned magic;           /* LIVE until fsm_b200_dfa_free; the struct itself is never
	                                    * returned to malloc so that a use after free is DETECTED */
	struct oracle_owned_desc copy;     /* deep copy of the description */
	struct fsm_b200_desc_ext ext;      /* the copy again, as an ext desc when it has eager outputs */
	uint32_t nbits;                    /* distinct eager-output ids, ascending */
	uint32_t *id_of_bit;
};

static unsigned long n_compile, n_free, n_stream, n_batch, n_det, n_min, n_uaf;   /* atomics */
#define COUNT(c) __atomic_fetch_add(&(c), 1, __ATOMIC_RELAXED)

/* FSM_B200_STUB_SLOW=<microseconds>: sleep inside every exec call, to widen the window in which
 * another thread can evict the table this call is using */
static void
maybe_stall(void)
{
	static int us = -1;
	if (us < 0) { const char *e = getenv("FSM_B200_STUB_SLOW"); us = e ? atoi(e) : 0; }
	if (us > 0) { struct timespec ts = { 0, (long) us * 1000L }; nanosleep(&ts, NULL); }
}

static void *
dup_block(const void *p, size_t bytes)
{
	void *q = malloc(bytes ? bytes : 1);
	if (q != NULL && bytes) memcpy(q, p, bytes);
	return q;
}

static int
cmp_u32(const void *a, const void *b)
{
	const uint32_t x = *(const uint32_t *) a, y = *(const uint32_t *) b;
	return x < y ? -1 : x > y;
}

const char *fsm_b200_last_error(void) { return "stub engine (CPU, tests only)"; }

int
fsm_b200_dfa_compile(const struct fsm_b200_desc *d, int device, fsm_b200_dfa **out)
{
	fsm_b200_dfa *dfa;
	const uint32_t n = d->nstates;
	uint64_t G;
	(void) device;
	COUNT(n_compile);
	if (!oracle_isdfa(d)) { errno = EINVAL; return -1; }      /* exec.c:106-114 */
	dfa = calloc(1, sizeof *dfa);
	if (dfa == NULL) { errno = ENOMEM; return -1; }
	G = d->group_off[n];
	dfa->magic = LIVE;
	dfa->copy.desc = *d;
	dfa->copy.desc.is_end = dfa->copy.blocks[0] = dup_block(d->is_end, n);
	dfa->copy.desc.group_off = dfa->copy.blocks[1] = dup_block(d->group_off, (n + 1) * sizeof(uint64_t));
	dfa->copy.desc.group_symbols = dfa->copy.blocks[2] = dup_block(d->group_symbols, 4 * G * sizeof(uint64_t));
	dfa->copy.desc.group_to = dfa->copy.blocks[3] = dup_block(d->group_to, G * sizeof(uint32_t));
	dfa->copy.desc.eps_off = NULL; dfa->copy.desc.eps_to = NULL;    /* a DFA has none */
	dfa->copy.desc.endid_off = NULL; dfa->copy.desc.endids = NULL;
	dfa->copy.desc.reserved = 0;
	dfa->ext.base = dfa->copy.desc;
	if (d->reserved & FSM_B200_DESC_EAGER) {
		const struct fsm_b200_desc_ext *x = (const struct fsm_b200_desc_ext *) d;
		const uint64_t total = x->eager_off[n];
		uint64_t i; uint32_t w = 0;
		dfa->ext.eager_off = dfa->copy.blocks[4] = dup_block(x->eager_off, (n + 1) * sizeof(uint64_t));
		dfa->ext.eager_ids = dfa->copy.blocks[5] = dup_block(x->eager_ids, total * sizeof(uint32_t));
		dfa->ext.base.reserved = FSM_B200_DESC_EAGER;
		dfa->id_of_bit = dup_block(x->eager_ids, total * sizeof(uint32_t));
		qsort(dfa->id_of_bit, total, sizeof(uint32_t), cmp_u32);
		for (i = 0; i < total; i++) if (i == 0 || dfa->id_of_bit[i] != dfa->id_of_bit[w - 1]) dfa->id_of_bit[w++] = dfa->id_of_bit[i];
		dfa->nbits = w;
		if (w > FSM_B200_EAGER_MAX_IDS) { fsm_b200_dfa_free(dfa); errno = ENOTSUP; return -1; }
	}
	*out = dfa;
	return 0;
}

void
fsm_b200_dfa_free(fsm_b200_dfa *dfa)
{
	int i;
	if (dfa == NULL) return;
	COUNT(n_free);
	if (dfa->magic != LIVE) { COUNT(n_uaf); return; }                 /* double free */
	dfa->magic = DEAD;
	for (i = 0; i < 8; i++) { free(dfa->copy.blocks[i]); dfa->copy.blocks[i] = NULL; }
	free(dfa->id_of_bit); dfa->id_of_bit = NULL; dfa->nbits = 0;
	memset(&dfa->copy.desc, 0, sizeof dfa->copy.desc);           /* header kept as a tombstone */
	memset(&dfa->ext, 0, sizeof dfa->ext);
}

int
fsm_b200_exec_stream_host(const fsm_b200_dfa *dfa, const uint8_t *buf, uint64_t len, struct fsm_b200_result *out)
{
	COUNT(n_stream);
	maybe_stall();
	if (dfa->magic != LIVE) { COUNT(n_uaf); errno = EFAULT; return -1; }
	oracle_exec(&dfa->copy.desc, buf, len, 0, out);
	if (dfa->magic != LIVE) { COUNT(n_uaf); errno = EFAULT; return -1; }
	return 0;
}

int
fsm_b200_exec_batch_host(const fsm_b200_dfa *dfa, const uint8_t *base, const uint64_t *offsets, size_t n,
	struct fsm_b200_result *out)
{
	COUNT(n_batch);
	maybe_stall();
	if (dfa->magic != LIVE) { COUNT(n_uaf); errno = EFAULT; return -1; }
	return oracle_exec_batch(&dfa->copy.desc, base, offsets, n, 0, 1, out);
}

static int
hand_over(struct oracle_owned_desc *od, struct fsm_b200_owned_desc *out)
{
	struct oracle_owned_desc *keep = malloc(sizeof *keep);
	if (keep == NULL) { oracle_desc_free(od); errno = ENOMEM; return -1; }
	*keep = *od;
	out->desc = keep->desc;
	out->owner = keep;
	return 0;
}

int
fsm_b200_dfa_eager_info(const fsm_b200_dfa *dfa, uint32_t *nbits, const uint32_t **id_of_bit)
{
	*nbits = dfa->nbits;
	*id_of_bit = dfa->id_of_bit;
	return 0;
}

int
fsm_b200_exec_batch_eager_host(const fsm_b200_dfa *dfa, const uint8_t *base, const uint64_t *offsets, size_t n,
	struct fsm_b200_result *out, uint64_t *masks)
{
	const size_t words = (dfa->nbits + 63u) / 64u;
	size_t i;
	COUNT(n_batch);
	maybe_stall();
	if (dfa->magic != LIVE) { COUNT(n_uaf); errno = EFAULT; return -1; }
	for (i = 0; i < n; i++) {
		uint32_t fired[FSM_B200_EAGER_MAX_IDS];
		size_t nf = 0, k;
		oracle_exec_eager(&dfa->ext.base, base + offsets[i], offsets[i + 1] - offsets[i], &out[i], fired,
		    FSM_B200_EAGER_MAX_IDS, &nf);
		memset(masks + i * words, 0, words * sizeof *masks);
		for (k = 0; k < nf; k++) {
			const uint32_t *p = bsearch(&fired[k], dfa->id_of_bit, dfa->nbits, sizeof(uint32_t), cmp_u32);
			const size_t b = (size_t) (p - dfa->id_of_bit);
			masks[i * words + (b >> 6)] |= 1ull << (b & 63);
		}
	}
	return 0;
}

int
fsm_b200_owned_desc_eager(const struct fsm_b200_owned_desc *d, const uint64_t **eager_off, const uint32_t **eager_ids)
{
	const struct oracle_owned_desc *keep = d->owner;
	*eager_off = keep != NULL ? keep->eager_off : NULL;
	*eager_ids = keep != NULL ? keep->eager_ids : NULL;
	return 0;
}

/* oracle/refnum_host.cpp: the product's refnum.h functions on the CPU */
int refnum_host_determinise_desc(const struct fsm_b200_desc *nfa, uint32_t state_limit, struct oracle_owned_desc *out);

int
fsm_b200_determinise_ex(const struct fsm_b200_desc *nfa, int device, size_t state_limit, unsigned flags,
	struct fsm_b200_owned_desc *out)
{
	struct oracle_owned_desc od;
	int rc;
	(void) device;
	COUNT(n_det);
	memset(out, 0, sizeof *out);
	if ((flags & FSM_B200_DET_REFERENCE_NUMBERING) && (nfa->reserved & FSM_B200_DESC_EAGER)) {
		errno = ENOTSUP;       /* the CPU harness does not combine the two; the engine does */
		return -1;
	}
	if (flags & FSM_B200_DET_REFERENCE_NUMBERING) {
		/* same verdicts as the engine: the input count is checked first (determinise.c:65-68),
		 * then at most limit+1 states may exist (determinise.c:166-169) */
		if (state_limit != 0 && nfa->nstates > state_limit) return 1;
		rc = refnum_host_determinise_desc(nfa, (uint32_t) state_limit, &od);
	} else {
		rc = oracle_determinise(nfa, state_limit, &od);
	}
	if (rc != 0) return rc;
	return hand_over(&od, out);
}

int
fsm_b200_determinise(const struct fsm_b200_desc *nfa, int device, size_t state_limit, struct fsm_b200_owned_desc *out)
{
	const char *e = getenv("FSM_B200_DET_NUMBERING");
	return fsm_b200_determinise_ex(nfa, device, state_limit,
	    (e != NULL && strcmp(e, "reference") == 0) ? FSM_B200_DET_REFERENCE_NUMBERING : 0u, out);
}

int
fsm_b200_minimise(const struct fsm_b200_desc *dfa, int device, struct fsm_b200_owned_desc *out)
{
	struct oracle_owned_desc od;
	(void) device;
	COUNT(n_min);
	memset(out, 0, sizeof *out);
	if (!oracle_isdfa(dfa)) { errno = EINVAL; return -1; }
	if (oracle_minimise(dfa, &od) != 0) return -1;
	return hand_over(&od, out);
}

void
fsm_b200_desc_free(struct fsm_b200_owned_desc *d)
{
	if (d == NULL || d->owner == NULL) return;
	oracle_desc_free(d->owner);
	free(d->owner);
	memset(d, 0, sizeof *d);
}

/* test hook: compile, free, stream, batch, determinise, minimise calls; uses after free */
void
fsm_b200_stub_counts(unsigned long out[7])
{
	out[0] = n_compile; out[1] = n_free; out[2] = n_stream; out[3] = n_batch; out[4] = n_det; out[5] = n_min;
	out[6] = n_uaf;
}
