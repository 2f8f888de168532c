/*
 * fsm_b200_shim.c -- libfsm's fsm_exec, re-implemented over libfsm_b200.so.
 *
 * Replaces src/libfsm/exec.c (and, further down, src/libfsm/determinise.c's two entry
 * points) of the reference.  What fsm_exec keeps, bit for bit:
 *   - signature and return convention: 1 match (+ *end), 0 no match, -1/errno
 *     (EINVAL when the fsm is not a DFA or has no start state, exec.c:106-114);
 *   - `*end` is written only on success (exec.c:165);
 *   - the input is read through the caller's getc callback; after a failed match on a
 *     missing edge the cursor of an fsm_sgetc string is left where the reference leaves it
 *     (one past the byte that had no edge, exec.c:132-138).  A FILE* is rewound the same way
 *     when it is seekable.
 * What it does differently, on purpose:
 *   - DFA-ness is validated once per (fsm, content fingerprint), not on every call: the
 *     reference's per-call fsm_all(fsm, fsm_isdfa) (exec.c:106) costs O(edges) per input;
 *   - the walk runs on the GPU (one long input: K1b chunk maps; a batch: K1).  There is no
 *     CPU walk in here: without a usable device the call fails with -1/EIO.
 * Eager outputs (exec.c:126-144): the engine returns the SET of ids fired along the walk; the
 * callback set with fsm_eager_output_set_cb is then called once per fired id, ascending, after
 * the walk (the reference calls it as it goes, once per state entry; its own tests only use the
 * set, tests/eager_output/utils.c:10-24).
 * Not accelerated (returns -1/ENOTSUP): FSMs with capture actions when `captures` is
 * non-NULL (exec.c:41-44,157-163) -- not produced by re_comp, no CLI passes captures.
 */
#include <assert.h>
#include <stddef.h>
#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fsm/fsm.h>
#include <fsm/capture.h>
#include <fsm/pred.h>
#include <fsm/walk.h>

#include <adt/set.h>
#include <adt/stateset.h>
#include <adt/edgeset.h>

#include "libfsm/internal.h"
#include "libfsm/eager_output.h"

#include "fsm_b200_shim.h"

/* ------------------------------------------------------------------ flattening ------- */

static int
cmp_u32(const void *a, const void *b)
{
	const uint32_t x = *(const uint32_t *) a, y = *(const uint32_t *) b;
	return x < y ? -1 : x > y;
}

/* struct fsm_b200_flat starts like a struct fsm_b200_desc_ext */
typedef char flat_is_ext[(offsetof(struct fsm_b200_flat, eager_off) == offsetof(struct fsm_b200_desc_ext, eager_off) &&
	offsetof(struct fsm_b200_flat, eager_ids) == offsetof(struct fsm_b200_desc_ext, eager_ids)) ? 1 : -1];

int
fsm_b200_flatten(const struct fsm *fsm, struct fsm_b200_flat *out)
{
	const size_t n = fsm->statecount;
	size_t ngroups = 0, neps = 0, nids = 0, nxids = 0, s;
	uint8_t *is_end; uint64_t *goff, *gsym, *eoff, *ioff, *xoff = NULL; uint32_t *gto, *eto, *ids, *xids = NULL;
	fsm_state_t start;

	memset(out, 0, sizeof *out);
	for (s = 0; s < n; s++) {
		struct edge_group_iter it;
		struct edge_group_iter_info info;
		edge_set_group_iter_reset(fsm->states[s].edges, EDGE_GROUP_ITER_ALL, &it);
		while (edge_set_group_iter_next(&it, &info)) ngroups++;
		neps += state_set_count(fsm->states[s].epsilons);
		if (fsm->states[s].end) nids += fsm_endid_count(fsm, (fsm_state_t) s);
		if (fsm->states[s].has_eager_outputs) nxids += fsm_eager_output_count(fsm, (fsm_state_t) s);
	}
	is_end = calloc(n + 1, 1);
	goff = calloc(n + 1, sizeof *goff);
	gsym = calloc(4 * ngroups + 4, sizeof *gsym);
	gto = calloc(ngroups + 1, sizeof *gto);
	eoff = calloc(n + 1, sizeof *eoff);
	eto = calloc(neps + 1, sizeof *eto);
	ioff = calloc(n + 1, sizeof *ioff);
	ids = calloc(nids + 1, sizeof *ids);
	if (nxids > 0) {
		xoff = calloc(n + 1, sizeof *xoff);
		xids = calloc(nxids + 1, sizeof *xids);
	}
	if (!is_end || !goff || !gsym || !gto || !eoff || !eto || !ioff || !ids || (nxids > 0 && (!xoff || !xids))) {
		free(is_end); free(goff); free(gsym); free(gto); free(eoff); free(eto); free(ioff); free(ids); free(xoff); free(xids);
		errno = ENOMEM;
		return -1;
	}
	ngroups = neps = nids = nxids = 0;
	for (s = 0; s < n; s++) {
		struct edge_group_iter it;
		struct edge_group_iter_info info;
		struct state_iter si;
		fsm_state_t es;

		goff[s] = ngroups; eoff[s] = neps; ioff[s] = nids;
		is_end[s] = (uint8_t) fsm->states[s].end;
		edge_set_group_iter_reset(fsm->states[s].edges, EDGE_GROUP_ITER_ALL, &it);
		while (edge_set_group_iter_next(&it, &info)) {
			memcpy(&gsym[4 * ngroups], info.symbols, sizeof info.symbols);
			gto[ngroups++] = info.to;
		}
		for (state_set_reset(fsm->states[s].epsilons, &si); state_set_next(&si, &es); ) {
			eto[neps++] = es;
		}
		if (fsm->states[s].end) {
			const size_t c = fsm_endid_count(fsm, (fsm_state_t) s);
			if (c > 0 && fsm_endid_get(fsm, (fsm_state_t) s, c, &ids[nids])) nids += c;
		}
		if (xoff != NULL) {
			/* eager outputs: stored in insertion order (eager_output.c:160-216); sets here */
			size_t c = fsm->states[s].has_eager_outputs ? fsm_eager_output_count(fsm, (fsm_state_t) s) : 0, i, w = 0;
			xoff[s] = nxids;
			if (c > 0) {
				fsm_eager_output_get(fsm, (fsm_state_t) s, c, &xids[nxids]);
				qsort(&xids[nxids], c, sizeof *xids, cmp_u32);
				for (i = 0; i < c; i++) if (i == 0 || xids[nxids + i] != xids[nxids + w - 1]) xids[nxids + w++] = xids[nxids + i];
				nxids += w;
			}
		}
	}
	goff[n] = ngroups; eoff[n] = neps; ioff[n] = nids;
	if (xoff != NULL) xoff[n] = nxids;
	out->desc.nstates = (uint32_t) n;
	out->desc.hasstart = (uint32_t) fsm_getstart(fsm, &start);
	out->desc.start = out->desc.hasstart ? start : 0;
	out->desc.is_end = is_end;
	out->desc.group_off = goff; out->desc.group_symbols = gsym; out->desc.group_to = gto;
	out->desc.eps_off = eoff; out->desc.eps_to = eto;
	out->desc.endid_off = ioff; out->desc.endids = ids;
	out->blocks[0] = is_end; out->blocks[1] = goff; out->blocks[2] = gsym; out->blocks[3] = gto;
	out->blocks[4] = eoff; out->blocks[5] = eto; out->blocks[6] = ioff; out->blocks[7] = ids;
	if (xoff != NULL) {
		out->desc.reserved = FSM_B200_DESC_EAGER;
		out->eager_off = xoff; out->eager_ids = xids;
		out->blocks[8] = xoff; out->blocks[9] = xids;
	}
	return 0;
}

void
fsm_b200_flat_free(struct fsm_b200_flat *flat)
{
	size_t i;
	for (i = 0; i < sizeof flat->blocks / sizeof flat->blocks[0]; i++) { free(flat->blocks[i]); flat->blocks[i] = NULL; }
	memset(&flat->desc, 0, sizeof flat->desc);
	flat->eager_off = NULL; flat->eager_ids = NULL;
}

#ifndef FSM_B200_SHIM_FLATTEN_ONLY

/* ------------------------------------------------------------------ device-table cache */

/* Content fingerprint: everything fsm_exec can observe (start, end bits, epsilon presence,
 * edge groups).  O(groups), allocation-free; the reference spends O(edges) per call on
 * validation alone. */
static int
fp_eager_cb(fsm_state_t state, fsm_output_id_t id, void *opaque)
{
	uint64_t *h = opaque;
	(void) state;
	*h += 0x9e3779b97f4a7c15ull * ((uint64_t) id + 1);     /* order-independent: stored order is insertion order */
	return 1;
}

static uint64_t
fingerprint(const struct fsm *fsm)
{
	uint64_t h = 0x9e3779b97f4a7c15ull ^ (uint64_t) fsm->statecount;
	size_t s;
#define MIX(v) do { h ^= (uint64_t) (v); h *= 0xff51afd7ed558ccdull; h ^= h >> 32; } while (0)
	MIX(fsm->hasstart); MIX(fsm->start);
	for (s = 0; s < fsm->statecount; s++) {
		struct edge_group_iter it;
		struct edge_group_iter_info info;
		MIX(fsm->states[s].end);
		MIX(state_set_count(fsm->states[s].epsilons));
		if (fsm->states[s].has_eager_outputs) {
			uint64_t e = 1;
			fsm_eager_output_iter_state(fsm, (fsm_state_t) s, fp_eager_cb, &e);
			MIX(e);
		}
		edge_set_group_iter_reset(fsm->states[s].edges, EDGE_GROUP_ITER_ALL, &it);
		while (edge_set_group_iter_next(&it, &info)) {
			MIX(info.to); MIX(info.symbols[0]); MIX(info.symbols[1]); MIX(info.symbols[2]); MIX(info.symbols[3]);
		}
		MIX(0x5bd1e995u);
	}
#undef MIX
	return h;
}

#define CACHE_SLOTS 16
static struct cache_entry {
	const struct fsm *fsm;
	uint64_t fp;
	fsm_b200_dfa *dfa;        /* NULL: known not to be a DFA (errno_val says why) */
	int errno_val;
	unsigned long stamp;      /* 0: slot unused */
	unsigned refs;            /* calls executing on `dfa` right now: not evictable */
	int doomed;               /* invalidated while in use: the last user frees it */
} cache[CACHE_SLOTS];
static unsigned long cache_clock;
static pthread_mutex_t cache_mu = PTHREAD_MUTEX_INITIALIZER;   /* lx(1) calls from threads */

/* A compiled table pinned for the duration of one call. */
struct dfa_ref {
	fsm_b200_dfa *dfa;
	struct cache_entry *entry;    /* NULL: not cached (every slot was in use), freed on release */
};

static int
device_index(void)
{
	const char *e = getenv("FSM_B200_DEVICE");
	return e != NULL ? atoi(e) : 0;
}

/* Pins the compiled DFA for `fsm` (cached) into *ref; 0, or -1 with errno set.
 * Release with put_dfa.  An entry is only ever evicted or freed while nobody holds it.
 * The flatten + validate + table build + upload of a miss run OUTSIDE cache_mu (lx(1) compiles
 * many zones from a thread pool): two threads that miss on the same automaton at the
 * same time both build it, the second one finds the first one's entry when it comes back and
 * drops its own copy. */
static int
get_dfa(const struct fsm *fsm, struct dfa_ref *ref)
{
	const uint64_t fp = fingerprint(fsm);
	struct cache_entry *victim = NULL;
	fsm_b200_dfa *dfa = NULL;
	struct fsm_b200_flat flat;
	int i, err = 0, pass;

	ref->dfa = NULL; ref->entry = NULL;
	for (pass = 0; pass < 2; pass++) {
		pthread_mutex_lock(&cache_mu);
		victim = NULL;
		for (i = 0; i < CACHE_SLOTS; i++) {
			struct cache_entry *e = &cache[i];
			if (e->stamp != 0 && !e->doomed && e->fsm == fsm && e->fp == fp) {
				e->stamp = ++cache_clock;
				if (e->dfa == NULL) {
					err = e->errno_val;
					pthread_mutex_unlock(&cache_mu);
					if (dfa != NULL) fsm_b200_dfa_free(dfa);
					errno = err;
					return -1;
				}
				e->refs++;
				ref->dfa = e->dfa; ref->entry = e;
				pthread_mutex_unlock(&cache_mu);
				if (dfa != NULL) fsm_b200_dfa_free(dfa);      /* somebody else was faster */
				return 0;
			}
			if (e->refs == 0 && !e->doomed && (victim == NULL || e->stamp < victim->stamp)) victim = e;
		}
		if (pass == 1) break;                /* still a miss, our build in hand: insert below, lock held */
		pthread_mutex_unlock(&cache_mu);

		/* miss: validate + build (the engine restates fsm_all(fsm_isdfa) + fsm_getstart) */
		if (fsm_b200_flatten(fsm, &flat) != 0) {
			return -1;
		}
		if (fsm_b200_dfa_compile(&flat.desc, device_index(), &dfa) != 0) {
			err = errno;
			dfa = NULL;
		}
		fsm_b200_flat_free(&flat);
	}
	if (victim != NULL && (dfa != NULL || err == EINVAL)) {      /* remember DFAs and definite non-DFAs */
		fsm_b200_dfa *old = (victim->stamp != 0) ? victim->dfa : NULL;
		memset(victim, 0, sizeof *victim);
		victim->fsm = fsm; victim->fp = fp; victim->dfa = dfa; victim->errno_val = err;
		victim->stamp = ++cache_clock;
		if (dfa != NULL) { victim->refs = 1; ref->entry = victim; }
		pthread_mutex_unlock(&cache_mu);
		if (old != NULL) fsm_b200_dfa_free(old);                 /* device frees outside the lock too */
	} else {
		pthread_mutex_unlock(&cache_mu);
	}
	if (dfa == NULL) {
		errno = err;
		return -1;
	}
	ref->dfa = dfa;
	return 0;
}

static void
put_dfa(struct dfa_ref *ref)
{
	const int saved = errno;
	if (ref->dfa == NULL) return;
	if (ref->entry == NULL) {
		fsm_b200_dfa_free(ref->dfa);             /* was never cached */
	} else {
		pthread_mutex_lock(&cache_mu);
		if (--ref->entry->refs == 0 && ref->entry->doomed) {
			fsm_b200_dfa_free(ref->entry->dfa);
			memset(ref->entry, 0, sizeof *ref->entry);
		}
		pthread_mutex_unlock(&cache_mu);
	}
	ref->dfa = NULL; ref->entry = NULL;
	errno = saved;
}

void
fsm_b200_invalidate(const struct fsm *fsm)
{
	int i;
	pthread_mutex_lock(&cache_mu);
	for (i = 0; i < CACHE_SLOTS; i++) {
		struct cache_entry *e = &cache[i];
		if (e->stamp == 0 || e->doomed || e->fsm != fsm) continue;
		if (e->refs > 0) {
			e->doomed = 1;                       /* in use on another thread: freed by its put_dfa */
			continue;
		}
		if (e->dfa != NULL) fsm_b200_dfa_free(e->dfa);
		memset(e, 0, sizeof *e);
	}
	pthread_mutex_unlock(&cache_mu);
}

static int
unsupported(const struct fsm *fsm, const struct fsm_capture *captures)
{
	return captures != NULL && fsm_countcaptures(fsm) > 0;
}

/* ------------------------------------------------------------------ fsm_exec ---------- */

int
fsm_exec(const struct fsm *fsm,
	int (*fsm_getc)(void *opaque), void *opaque,
	fsm_state_t *end, struct fsm_capture *captures)
{
	struct dfa_ref ref;
	struct fsm_b200_result r;
	unsigned char *buf = NULL;
	const unsigned char *data = NULL;
	size_t len = 0, cap = 0;
	const char *sgetc_start = NULL;
	long file_start = -1;
	int c;

	assert(fsm != NULL);
	assert(fsm_getc != NULL);
	assert(end != NULL);

	if (unsupported(fsm, captures)) {
		errno = ENOTSUP;
		return -1;
	}
	if (get_dfa(fsm, &ref) != 0) {      /* -1/EINVAL: not a DFA, no start (exec.c:106-114) */
		return -1;
	}

	if (fsm_getc == fsm_sgetc) {
		/* the reference's own string cursor (getc.c:16-34): the bytes are already in memory, up to the
		 * NUL -- take them in place instead of one indirect call per byte, and leave the cursor where
		 * draining would have left it (on the NUL; fsm_sgetc does not step past it) */
		sgetc_start = *(const char **) opaque;
		len = strlen(sgetc_start);
		data = (const unsigned char *) sgetc_start;
		*(const char **) opaque = sgetc_start + len;
	} else if (fsm_getc == fsm_fgetc) {
		/* the reference's own FILE cursor (getc.c:36-51): block reads instead of fgetc per byte */
		FILE *f = opaque;
		file_start = ftell(f);
		for (;;) {
			size_t got;
			if (cap - len < (1u << 20)) {
				size_t ncap = cap ? cap * 2 : (4u << 20);
				unsigned char *nb = realloc(buf, ncap);
				if (nb == NULL) {
					free(buf);
					put_dfa(&ref);
					errno = ENOMEM;
					return -1;
				}
				buf = nb; cap = ncap;
			}
			got = fread(buf + len, 1, cap - len, f);
			len += got;
			if (got == 0) break;          /* EOF or error: fsm_fgetc returns EOF for both */
		}
		data = buf;
	} else {
		/* any other callback: drain it (exec.c:132) */
		while (c = fsm_getc(opaque), c != EOF) {
			if (len == cap) {
				size_t ncap = cap ? cap * 2 : 4096;
				unsigned char *nb = realloc(buf, ncap);
				if (nb == NULL) {
					free(buf);
					put_dfa(&ref);
					errno = ENOMEM;
					return -1;
				}
				buf = nb; cap = ncap;
			}
			buf[len++] = (unsigned char) c;
		}
		data = buf;
	}

	{
		uint32_t nbits = 0, b;
		const uint32_t *id_of_bit = NULL;
		uint64_t mask[FSM_B200_EAGER_MAX_IDS / 64];
		const uint64_t off[2] = { 0, len };
		int rc;

		(void) fsm_b200_dfa_eager_info(ref.dfa, &nbits, &id_of_bit);
		if (nbits == 0) {
			rc = fsm_b200_exec_stream_host(ref.dfa, data, len, &r);
		} else {
			/* eager outputs (exec.c:126-144): the walk also reports the set of ids it fired */
			memset(mask, 0, sizeof mask);
			rc = fsm_b200_exec_batch_eager_host(ref.dfa, data != NULL ? data : (const unsigned char *) "", off, 1, &r, mask);
		}
		if (rc != 0) {
			free(buf);
			put_dfa(&ref);
			return -1;                  /* errno from the engine (EIO: no device) */
		}
		if (nbits != 0) {
			fsm_eager_output_cb *cb = NULL;
			void *cb_opaque = NULL;
			fsm_eager_output_get_cb(fsm, &cb, &cb_opaque);
			for (b = 0; cb != NULL && b < nbits; b++) {
				if ((mask[b >> 6] >> (b & 63)) & 1u) cb(id_of_bit[b], cb_opaque);
			}
		}
	}
	free(buf);
	put_dfa(&ref);

	if (r.ret == 0 && r.consumed < len) {
		/* the reference stopped reading right after the byte with no edge (exec.c:133-138) */
		if (sgetc_start != NULL) {
			*(const char **) opaque = sgetc_start + r.consumed + 1;
		} else if (file_start >= 0) {
			(void) fseek((FILE *) opaque, file_start + (long) r.consumed + 1, SEEK_SET);
		}
	}
	if (r.ret != 1) {
		return 0;
	}
	*end = r.end;
	return 1;
}

/* Build a `struct fsm` from a flat DFA description with the reference's own builder API.
 * Groups arrive sorted by destination, so every edge_set_add_bulk is a pure append. */
static struct fsm *
fsm_from_desc(const struct fsm_alloc *alloc, const struct fsm_b200_desc *d,
	const uint64_t *eager_off, const uint32_t *eager_ids)
{
	struct fsm *fsm = fsm_new_statealloc(alloc, d->nstates > 0 ? d->nstates : 1);
	uint32_t s;

	if (fsm == NULL) {
		return NULL;
	}
	if (d->nstates > 0 && !fsm_addstate_bulk(fsm, d->nstates)) {
		goto fail;
	}
	for (s = 0; s < d->nstates; s++) {
		uint64_t g, e;
		const uint64_t g0 = d->group_off[s], g1 = d->group_off[s + 1];
		if (g1 > g0 && !edge_set_advise_growth(&fsm->states[s].edges, fsm->alloc, (size_t) (g1 - g0))) {
			goto fail;
		}
		for (g = g0; g < g1; g++) {
			uint64_t sym[4];
			memcpy(sym, &d->group_symbols[4 * g], sizeof sym);
			if (!edge_set_add_bulk(&fsm->states[s].edges, fsm->alloc, sym, d->group_to[g])) {
				goto fail;
			}
		}
		if (d->is_end[s]) {
			fsm_setend(fsm, s, 1);
			for (e = d->endid_off[s]; e < d->endid_off[s + 1]; e++) {
				if (!fsm_endid_set(fsm, s, d->endids[e])) {
					goto fail;
				}
			}
		}
	}
	if (d->hasstart) {
		fsm_setstart(fsm, d->start);
	}
	if (eager_off != NULL) {
		for (s = 0; s < d->nstates; s++) {
			uint64_t e;
			for (e = eager_off[s]; e < eager_off[s + 1]; e++) {
				if (!fsm_eager_output_set(fsm, s, eager_ids[e])) {
					goto fail;
				}
			}
		}
	}
	return fsm;

fail:
	fsm_free(fsm);
	return NULL;
}

/* ------------------------------------------------------------------ fsm_determinise --- */

/*
 * fsm_determinise_with_config / fsm_determinise (include/fsm/fsm.h:472-488), replacing
 * src/libfsm/determinise.c.  Same contract: in place -- afterwards the same `struct fsm *`
 * holds the DFA, state 0 is the start state (determinise.c:234), the allocator is kept,
 * end bits and end ids are carried (determinise.c:236-266), linkage_info is gone
 * (determinise.c:284).  The subset construction itself (incl. the epsilon removal the
 * reference runs first, determinise.c:48-52) happens in the engine (K2); this function
 * only marshals: struct fsm -> flat NFA -> [GPU] -> flat DFA -> struct fsm -> fsm_move.
 * State NUMBERS are BFS order instead of the reference's LIFO/analysis order (DESIGN.md
 * section 5): an isomorphic DFA -- unless FSM_B200_DET_NUMBERING=reference, in which case the
 * engine reproduces the reference's numbering too.
 * Eager outputs are carried by the engine (determinise.c:268-274, :2614-2636).
 * Not accelerated: capture actions (determinise.c:268-270 remaps them); such FSMs fail with
 * ERRNO/ENOTSUP rather than silently taking another path.
 */
enum fsm_determinise_with_config_res
fsm_determinise_with_config(struct fsm *nfa, const struct fsm_determinise_config *config)
{
	const size_t state_limit = config == NULL ? 0 : config->state_limit;
	struct fsm_b200_flat flat;
	struct fsm_b200_owned_desc out;
	struct fsm *dfa;
	int rc;

	assert(nfa != NULL);
	if (fsm_countcaptures(nfa) > 0 || unsupported(nfa, NULL)) {
		errno = ENOTSUP;
		return FSM_DETERMINISE_WITH_CONFIG_ERRNO;
	}
	if (fsm_b200_flatten(nfa, &flat) != 0) {
		return FSM_DETERMINISE_WITH_CONFIG_ERRNO;
	}
	rc = fsm_b200_determinise(&flat.desc, device_index(), state_limit, &out);
	if (rc == 1) {
		fsm_b200_flat_free(&flat);
		return FSM_DETERMINISE_WITH_CONFIG_STATE_LIMIT_REACHED;
	}
	if (rc != 0) {
		fsm_b200_flat_free(&flat);
		return FSM_DETERMINISE_WITH_CONFIG_ERRNO;
	}
	if (!flat.desc.hasstart) {
		/* determinise.c:88-91: no start state => OK, and the fsm only loses its epsilons */
		fsm_b200_flat_free(&flat);
		fsm_b200_desc_free(&out);
		if (fsm_has(nfa, fsm_hasepsilons) && !fsm_remove_epsilons(nfa)) {
			return FSM_DETERMINISE_WITH_CONFIG_ERRNO;
		}
		return FSM_DETERMINISE_WITH_CONFIG_OK;
	}
	fsm_b200_flat_free(&flat);

	{
		/* eager outputs carried by the engine (determinise.c:268-274); like the reference, the
		 * callback registered on the old automaton does not survive the fsm_move */
		const uint64_t *xoff = NULL; const uint32_t *xids = NULL;
		(void) fsm_b200_owned_desc_eager(&out, &xoff, &xids);
		dfa = fsm_from_desc(nfa->alloc, &out.desc, xoff, xids);
		fsm_b200_desc_free(&out);
		if (dfa == NULL) {
			return FSM_DETERMINISE_WITH_CONFIG_ERRNO;
		}
	}
	fsm_b200_invalidate(nfa);
	fsm_move(nfa, dfa);
	return FSM_DETERMINISE_WITH_CONFIG_OK;
}

int
fsm_determinise(struct fsm *nfa)
{
	switch (fsm_determinise_with_config(nfa, NULL)) {
	case FSM_DETERMINISE_WITH_CONFIG_OK:
		return 1;
	case FSM_DETERMINISE_WITH_CONFIG_STATE_LIMIT_REACHED:
	case FSM_DETERMINISE_WITH_CONFIG_ERRNO:
	default:
		return 0;
	}
}

/* ------------------------------------------------------------------ fsm_minimise ------ */

/*
 * fsm_minimise (include/fsm/fsm.h:502-503), replacing src/libfsm/minimise.c's entry point.
 * Same contract: requires a DFA (the reference asserts, minimise.c:89-90; here 0/EINVAL), in
 * place, 1 on success.  Trim + partition refinement run in the engine (K3); the minimal DFA is
 * unique up to numbering, so the result is isomorphic to the reference's.  An fsm that can
 * match nothing ends up with no states, like the reference after its fsm_trim
 * (minimise.c:93-101).
 */
int
fsm_minimise(struct fsm *fsm)
{
	struct fsm_b200_flat flat;
	struct fsm_b200_owned_desc out;
	struct fsm *min;

	assert(fsm != NULL);
	if (fsm_countcaptures(fsm) > 0 || unsupported(fsm, NULL)) {
		errno = ENOTSUP;
		return 0;
	}
	if (fsm->statecount == 0) {
		return 1;
	}
	if (fsm_b200_flatten(fsm, &flat) != 0) {
		return 0;
	}
	if (!flat.desc.hasstart) {           /* fsm_trim: nothing is reachable; the reference sweeps every state */
		fsm_b200_flat_free(&flat);
		min = fsm_new(fsm->alloc);
		if (min == NULL) return 0;
		fsm_b200_invalidate(fsm);
		fsm_move(fsm, min);
		return 1;
	}
	if (fsm_b200_minimise(&flat.desc, device_index(), &out) != 0) {
		fsm_b200_flat_free(&flat);
		return 0;
	}
	fsm_b200_flat_free(&flat);
	{
		const uint64_t *xoff = NULL; const uint32_t *xids = NULL;
		(void) fsm_b200_owned_desc_eager(&out, &xoff, &xids);
		min = fsm_from_desc(fsm->alloc, &out.desc, xoff, xids);
		fsm_b200_desc_free(&out);
		if (min == NULL) {
			return 0;
		}
	}
	fsm_b200_invalidate(fsm);
	fsm_move(fsm, min);
	return 1;
}

int
fsm_exec_batch(const struct fsm *fsm, const unsigned char *base, const uint64_t *offsets,
	size_t n, struct fsm_b200_result *out)
{
	struct dfa_ref ref;
	int rc;

	assert(fsm != NULL);
	if (unsupported(fsm, NULL)) {
		errno = ENOTSUP;
		return -1;
	}
	if (get_dfa(fsm, &ref) != 0) {
		return -1;
	}
	rc = fsm_b200_exec_batch_host(ref.dfa, base, offsets, n, out);
	put_dfa(&ref);
	return rc;
}

int
fsm_exec_batch_eager(const struct fsm *fsm, const unsigned char *base, const uint64_t *offsets,
	size_t n, struct fsm_b200_result *out, uint64_t *masks, uint32_t *nbits, uint32_t *id_of_bit)
{
	struct dfa_ref ref;
	const uint32_t *ids = NULL;
	int rc;

	assert(fsm != NULL);
	assert(nbits != NULL && id_of_bit != NULL);
	if (get_dfa(fsm, &ref) != 0) {
		return -1;
	}
	rc = fsm_b200_dfa_eager_info(ref.dfa, nbits, &ids);
	if (rc == 0 && *nbits > 0) {
		memcpy(id_of_bit, ids, *nbits * sizeof *id_of_bit);      /* a copy: the cached table may be evicted after put_dfa */
	}
	if (rc == 0 && n > 0) {
		rc = *nbits != 0 ? fsm_b200_exec_batch_eager_host(ref.dfa, base, offsets, n, out, masks)
		                 : fsm_b200_exec_batch_host(ref.dfa, base, offsets, n, out);
	}
	put_dfa(&ref);
	return rc;
}

#endif /* FSM_B200_SHIM_FLATTEN_ONLY */

