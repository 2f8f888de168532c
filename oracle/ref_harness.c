/*
 * ref_harness.c -- TEST INFRASTRUCTURE.  See ref_harness.h.
 *
 * Compiled against the reference's headers where they lie (/root/reference/include and
 * its private src/libfsm/internal.h, include/adt/edgeset.h, include/adt/stateset.h) and
 * linked to oracle/_ref/libfsm_ref.so.  All automaton logic here is the reference's own;
 * this file only marshals between `struct fsm` and the flat description.
 */
#define _POSIX_C_SOURCE 200809L
#include <assert.h>
#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <fsm/fsm.h>
#include <fsm/bool.h>
#include <fsm/pred.h>
#include <fsm/walk.h>
#include <fsm/parser.h>
#include <re/re.h>
#include <fsm/vm.h>
#include <fsm/options.h>

#include <adt/set.h>
#include <adt/stateset.h>
#include <adt/edgeset.h>

#include "libfsm/internal.h"
#include "libfsm/eager_output.h"

#include "ref_harness.h"

struct strcur { const char *p; size_t left; };

static int
str_getc(void *opaque)
{
	struct strcur *c = opaque;
	if (c->left == 0) return EOF;
	c->left--;
	return (unsigned char) *c->p++;
}

void *
refh_re_comp(const char *pattern, size_t len, int dialect, int flags)
{
	struct strcur c = { pattern, len };
	struct re_err err;
	return re_comp((enum re_dialect) dialect, str_getc, &c, NULL, (enum re_flags) flags, &err);
}

void *
refh_parse_file(const char *path)
{
	FILE *f = fopen(path, "r");
	struct fsm *fsm;
	if (f == NULL) return NULL;
	fsm = fsm_parse(f, NULL);
	fclose(f);
	return fsm;
}

int refh_determinise(void *fsm) { return fsm_determinise(fsm); }

int
refh_determinise_limit(void *fsm, size_t state_limit)
{
	struct fsm_determinise_config cfg = { state_limit };
	return (int) fsm_determinise_with_config(fsm, &cfg);
}

int refh_minimise(void *fsm) { return fsm_minimise(fsm); }
int refh_setendid(void *fsm, unsigned id) { return fsm_setendid(fsm, id); }
void *refh_union_array(size_t n, void **fsms) { return fsm_union_array(n, (struct fsm **) fsms, NULL); }
void *refh_clone(const void *fsm) { return fsm_clone(fsm); }
void refh_free(void *fsm) { fsm_free(fsm); }
unsigned refh_countstates(const void *fsm) { return fsm_countstates(fsm); }
int refh_equal(const void *a, const void *b) { return fsm_equal(a, b); }
int refh_remove_epsilons(void *fsm) { return fsm_remove_epsilons(fsm); }

void *
refh_from_desc(const struct fsm_b200_desc *d)
{
	struct fsm *fsm = fsm_new(NULL);
	uint32_t s;
	if (fsm == NULL) return NULL;
	if (d->nstates > 0 && !fsm_addstate_bulk(fsm, d->nstates)) goto fail;
	for (s = 0; s < d->nstates; s++) {
		uint64_t g, e;
		for (g = d->group_off[s]; g < d->group_off[s + 1]; g++) {
			unsigned c;
			for (c = 0; c < 256; c++) {
				if ((d->group_symbols[4 * g + (c >> 6)] >> (c & 63)) & 1u) {
					if (!fsm_addedge_literal(fsm, s, d->group_to[g], (char) c)) goto fail;
				}
			}
		}
		if (d->eps_off != NULL) {
			for (e = d->eps_off[s]; e < d->eps_off[s + 1]; e++) {
				if (!fsm_addedge_epsilon(fsm, s, d->eps_to[e])) goto fail;
			}
		}
		if (d->is_end[s]) {
			fsm_setend(fsm, s, 1);
			if (d->endid_off != NULL) {
				for (e = d->endid_off[s]; e < d->endid_off[s + 1]; e++) {
					if (!fsm_endid_set(fsm, s, d->endids[e])) goto fail;
				}
			}
		}
	}
	if (d->hasstart) fsm_setstart(fsm, d->start);
	if (d->reserved & FSM_B200_DESC_EAGER) {
		const struct fsm_b200_desc_ext *x = (const struct fsm_b200_desc_ext *) d;
		for (s = 0; s < d->nstates; s++) {
			uint64_t e;
			for (e = x->eager_off[s]; e < x->eager_off[s + 1]; e++) {
				if (!fsm_eager_output_set(fsm, s, x->eager_ids[e])) goto fail;
			}
		}
	}
	return fsm;
fail:
	fsm_free(fsm);
	return NULL;
}

int
refh_flatten(const void *vfsm, struct refh_flat *out)
{
	const struct fsm *fsm = vfsm;
	const size_t n = fsm->statecount;
	size_t ngroups = 0, neps = 0, nids = 0, s;
	uint8_t *is_end; uint64_t *goff, *gsym, *eoff, *ioff; uint32_t *gto, *eto, *ids;
	fsm_state_t st;

	memset(out, 0, sizeof *out);
	for (s = 0; s < n; s++) {
		struct edge_group_iter egi;
		struct edge_group_iter_info info;
		edge_set_group_iter_reset(fsm->states[s].edges, EDGE_GROUP_ITER_ALL, &egi);
		while (edge_set_group_iter_next(&egi, &info)) ngroups++;
		neps += state_set_count(fsm->states[s].epsilons);
		if (fsm->states[s].end) nids += fsm_endid_count(fsm, (fsm_state_t) s);
	}
	is_end = calloc(n + 1, 1);
	goff = calloc(n + 1, sizeof *goff);
	gsym = calloc(4 * ngroups + 4, sizeof *gsym);
	gto  = calloc(ngroups + 1, sizeof *gto);
	eoff = calloc(n + 1, sizeof *eoff);
	eto  = calloc(neps + 1, sizeof *eto);
	ioff = calloc(n + 1, sizeof *ioff);
	ids  = calloc(nids + 1, sizeof *ids);
	if (!is_end || !goff || !gsym || !gto || !eoff || !eto || !ioff || !ids) {
		free(is_end); free(goff); free(gsym); free(gto); free(eoff); free(eto); free(ioff); free(ids);
		errno = ENOMEM;
		return -1;
	}
	ngroups = neps = nids = 0;
	for (s = 0; s < n; s++) {
		struct edge_group_iter egi;
		struct edge_group_iter_info info;
		struct state_iter si;
		fsm_state_t es;

		goff[s] = ngroups; eoff[s] = neps; ioff[s] = nids;
		is_end[s] = (uint8_t) fsm->states[s].end;
		edge_set_group_iter_reset(fsm->states[s].edges, EDGE_GROUP_ITER_ALL, &egi);
		while (edge_set_group_iter_next(&egi, &info)) {
			memcpy(&gsym[4 * ngroups], info.symbols, 4 * sizeof(uint64_t));
			gto[ngroups] = info.to;
			ngroups++;
		}
		for (state_set_reset(fsm->states[s].epsilons, &si); state_set_next(&si, &es); ) {
			eto[neps++] = es;
		}
		if (fsm->states[s].end) {
			size_t c = fsm_endid_count(fsm, (fsm_state_t) s);
			if (c > 0) {
				int ok = fsm_endid_get(fsm, (fsm_state_t) s, c, &ids[nids]);
				assert(ok); (void) ok;
				nids += c;
			}
		}
	}
	goff[n] = ngroups; eoff[n] = neps; ioff[n] = nids;

	out->desc.nstates = (uint32_t) n;
	out->desc.hasstart = (uint32_t) fsm_getstart(fsm, &st);
	out->desc.start = out->desc.hasstart ? st : 0;
	out->desc.is_end = is_end;
	out->desc.group_off = goff; out->desc.group_symbols = gsym; out->desc.group_to = gto;
	out->desc.eps_off = eoff; out->desc.eps_to = eto;
	out->desc.endid_off = ioff; out->desc.endids = ids;
	out->blocks[0] = is_end; out->blocks[1] = goff; out->blocks[2] = gsym; out->blocks[3] = gto;
	out->blocks[4] = eoff; out->blocks[5] = eto; out->blocks[6] = ioff; out->blocks[7] = ids;
	return 0;
}

void
refh_flat_free(struct refh_flat *f)
{
	size_t i;
	for (i = 0; i < 8; i++) { free(f->blocks[i]); f->blocks[i] = NULL; }
	memset(&f->desc, 0, sizeof f->desc);
}

int
refh_epsilon_closure(void *vfsm, uint64_t **off_out, uint32_t **to_out)
{
	struct fsm *fsm = vfsm;
	const size_t n = fsm->statecount;
	struct state_set **cl = epsilon_closure(fsm);
	uint64_t *off; uint32_t *to; size_t total = 0, s;
	if (cl == NULL) return -1;
	for (s = 0; s < n; s++) total += state_set_count(cl[s]);
	off = calloc(n + 1, sizeof *off);
	to = calloc(total + 1, sizeof *to);
	if (!off || !to) { free(off); free(to); closure_free(fsm, cl, n); errno = ENOMEM; return -1; }
	total = 0;
	for (s = 0; s < n; s++) {
		struct state_iter si; fsm_state_t es;
		off[s] = total;
		for (state_set_reset(cl[s], &si); state_set_next(&si, &es); ) to[total++] = es;
	}
	off[n] = total;
	closure_free(fsm, cl, n);
	*off_out = off; *to_out = to;
	return 0;
}

struct bufcur { const uint8_t *p; uint64_t len, nread; int hit_eof; };

static int
buf_getc(void *opaque)
{
	struct bufcur *c = opaque;
	if (c->nread == c->len) { c->hit_eof = 1; return EOF; }
	return c->p[c->nread++];
}

int
refh_exec(const void *fsm, const uint8_t *buf, uint64_t len, struct fsm_b200_result *out)
{
	struct bufcur c = { buf, len, 0, 0 };
	fsm_state_t end = (fsm_state_t) -1;
	int r = fsm_exec(fsm, buf_getc, &c, &end, NULL);
	out->ret = r;
	out->end = (r == 1) ? end : UINT32_MAX;
	/* offset at return == bytes successfully transitioned: every byte read, unless the
	 * walk stopped on a byte without an edge (that byte was read but not consumed) */
	out->consumed = c.hit_eof ? c.nread : (c.nread > 0 ? c.nread - 1 : 0);
	if (r < 0) out->consumed = 0;
	return r;
}

struct job {
	const struct fsm *fsm; const uint8_t *base; const uint64_t *offsets;
	size_t lo, hi; int mode; struct fsm_b200_result *out;
};

static void *
worker(void *opaque)
{
	struct job *j = opaque;
	size_t i;
	for (i = j->lo; i < j->hi; i++) {
		const uint8_t *buf = j->base + j->offsets[i];
		const uint64_t len = j->offsets[i + 1] - j->offsets[i];
		if (j->mode == 0) {
			refh_exec(j->fsm, buf, len, &j->out[i]);
		} else {
			/* validation hoisted; the reference's own per-byte transition */
			fsm_state_t state = j->fsm->start;
			uint64_t off = 0;
			int dead = 0;
			while (off < len) {
				fsm_state_t next;
				if (!edge_set_transition(j->fsm->states[state].edges, buf[off], &next)) {
					dead = 1;
					break;
				}
				state = next;
				off++;
			}
			j->out[i].ret = (!dead && fsm_isend(j->fsm, state)) ? 1 : 0;
			j->out[i].end = state;
			j->out[i].consumed = off;
		}
	}
	return NULL;
}

int
refh_exec_batch(const void *vfsm, const uint8_t *base, const uint64_t *offsets, size_t n,
	int mode, int nthreads, struct fsm_b200_result *out)
{
	const struct fsm *fsm = vfsm;
	pthread_t tids[256];
	struct job jobs[256];
	int t;
	fsm_state_t st;

	if (mode == 1) {
		if (!fsm_all(fsm, fsm_isdfa) || !fsm_getstart(fsm, &st)) { errno = EINVAL; return -1; }
	}
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 256) nthreads = 256;
	for (t = 0; t < nthreads; t++) {
		jobs[t].fsm = fsm; jobs[t].base = base; jobs[t].offsets = offsets;
		jobs[t].lo = n * (size_t) t / (size_t) nthreads;
		jobs[t].hi = n * (size_t) (t + 1) / (size_t) nthreads;
		jobs[t].mode = mode; jobs[t].out = out;
	}
	if (nthreads == 1) {
		worker(&jobs[0]);
		return 0;
	}
	for (t = 0; t < nthreads; t++) {
		if (pthread_create(&tids[t], NULL, worker, &jobs[t]) != 0) {
			int k;
			for (k = 0; k < t; k++) pthread_join(tids[k], NULL);
			errno = EAGAIN;
			return -1;
		}
	}
	for (t = 0; t < nthreads; t++) pthread_join(tids[t], NULL);
	return 0;
}

size_t
refh_endids(const void *fsm, unsigned state, unsigned *ids, size_t cap)
{
	size_t c;
	if (!fsm_isend(fsm, state)) return 0;
	c = fsm_endid_count(fsm, state);
	if (c > 0 && c <= cap) {
		int ok = fsm_endid_get(fsm, state, c, ids);
		(void) ok;
	}
	return c;
}

/* ---- eager outputs (include/fsm/fsm.h:273-336, src/libfsm/eager_output.c) ---------------- */

static int
cmp_u32(const void *a, const void *b)
{
	const unsigned x = *(const unsigned *) a, y = *(const unsigned *) b;
	return x < y ? -1 : x > y;
}

int
refh_eager_set(void *fsm, unsigned state, unsigned id)
{
	return fsm_eager_output_set(fsm, state, id);
}

int
refh_eager_flatten(const void *vfsm, uint64_t **off_out, uint32_t **ids_out)
{
	const struct fsm *fsm = vfsm;
	const size_t n = fsm->statecount;
	uint64_t *off = calloc(n + 1, sizeof *off);
	uint32_t *ids;
	size_t total = 0, s;
	if (off == NULL) return -1;
	for (s = 0; s < n; s++) total += fsm_eager_output_count(fsm, (fsm_state_t) s);
	ids = calloc(total + 1, sizeof *ids);
	if (ids == NULL) { free(off); return -1; }
	total = 0;
	for (s = 0; s < n; s++) {
		const size_t c = fsm_eager_output_count(fsm, (fsm_state_t) s);
		size_t i, w;
		off[s] = total;
		if (c == 0) continue;
		fsm_eager_output_get(fsm, (fsm_state_t) s, c, &ids[total]);
		qsort(&ids[total], c, sizeof *ids, cmp_u32);
		for (i = 0, w = 0; i < c; i++) if (i == 0 || ids[total + i] != ids[total + w - 1]) ids[total + w++] = ids[total + i];
		total += w;
	}
	off[n] = total;
	*off_out = off; *ids_out = ids;
	return 0;
}

struct fired { unsigned *ids; size_t cap, used; };

static void
fired_cb(fsm_output_id_t id, void *opaque)
{
	struct fired *f = opaque;
	size_t i;
	for (i = 0; i < f->used; i++) if (f->ids[i] == id) return;
	if (f->used < f->cap) f->ids[f->used] = id;
	f->used++;
}

/* One fsm_exec with the eager-output callback installed: the SET of ids the reference fires
 * (exec.c:126-144), ascending, whether or not the input matches.  *nfired may exceed cap. */
int
refh_exec_eager(void *vfsm, const uint8_t *buf, uint64_t len, struct fsm_b200_result *out,
	unsigned *fired, size_t cap, size_t *nfired)
{
	struct fsm *fsm = vfsm;
	struct fired f = { fired, cap, 0 };
	int r;
	fsm_eager_output_set_cb(fsm, fired_cb, &f);
	r = refh_exec(fsm, buf, len, out);
	fsm_eager_output_set_cb(fsm, NULL, NULL);
	qsort(fired, f.used < cap ? f.used : cap, sizeof *fired, cmp_u32);
	*nfired = f.used;
	return r;
}

/* fsm_union_repeated_pattern_group (include/fsm/bool.h): how the reference itself attaches
 * eager outputs to a union of unanchored patterns.  Consumes the inputs. */
void *
refh_union_repeated_pattern_group(size_t n, void **fsms, unsigned id_base)
{
	return fsm_union_repeated_pattern_group(n, (struct fsm **) fsms, NULL, id_base);
}

/* n fsm_exec calls with the eager-output callback, over nthreads pthreads; the callback slot lives
 * in the fsm (eager_output.c:27-63), so every thread works on its own fsm_clone.  masks[i*words..]
 * gets bit b set when id_of_bit[b] (ascending) fired on line i.
 * mode 0 "as-is": the reference's fsm_exec per line.  mode 1 "amortised": validation hoisted; per
 * byte the reference's own edge_set_transition, per state entered fsm_eager_output_iter_state. */
struct ejob {
	struct fsm *fsm; const uint8_t *base; const uint64_t *offsets;
	size_t lo, hi; int mode; struct fsm_b200_result *out;
	uint64_t *masks; size_t words; const uint32_t *id_of_bit; size_t nbits;
	uint64_t *cur;
};

static void
ejob_cb(fsm_output_id_t id, void *opaque)
{
	struct ejob *j = opaque;
	size_t lo = 0, hi = j->nbits;
	while (lo < hi) {
		const size_t mid = (lo + hi) / 2;
		if (j->id_of_bit[mid] < id) lo = mid + 1; else hi = mid;
	}
	if (lo < j->nbits && j->id_of_bit[lo] == id) j->cur[lo >> 6] |= (uint64_t) 1 << (lo & 63);
}

static int
ejob_iter_cb(fsm_state_t state, fsm_output_id_t id, void *opaque)
{
	(void) state;
	ejob_cb(id, opaque);
	return 1;
}

static void *
eworker(void *opaque)
{
	struct ejob *j = opaque;
	size_t i;
	if (j->mode == 0) fsm_eager_output_set_cb(j->fsm, ejob_cb, j);
	for (i = j->lo; i < j->hi; i++) {
		const uint8_t *buf = j->base + j->offsets[i];
		const uint64_t len = j->offsets[i + 1] - j->offsets[i];
		j->cur = j->masks + i * j->words;
		memset(j->cur, 0, j->words * sizeof *j->cur);
		if (j->mode == 0) {
			refh_exec(j->fsm, buf, len, &j->out[i]);
			if (j->out[i].ret != 1) {
				/* fsm_exec does not report where a failed walk stopped; the record's `end` does */
				fsm_state_t state = j->fsm->start, next;
				uint64_t off = 0;
				while (off < len && edge_set_transition(j->fsm->states[state].edges, buf[off], &next)) { state = next; off++; }
				j->out[i].end = state;
			}
		} else {
			fsm_state_t state = j->fsm->start;
			uint64_t off = 0;
			int dead = 0;
			if (j->fsm->states[state].has_eager_outputs) fsm_eager_output_iter_state(j->fsm, state, ejob_iter_cb, j);
			while (off < len) {
				fsm_state_t next;
				if (!edge_set_transition(j->fsm->states[state].edges, buf[off], &next)) { dead = 1; break; }
				state = next;
				if (j->fsm->states[state].has_eager_outputs) fsm_eager_output_iter_state(j->fsm, state, ejob_iter_cb, j);
				off++;
			}
			j->out[i].ret = (!dead && fsm_isend(j->fsm, state)) ? 1 : 0;
			j->out[i].end = state;
			j->out[i].consumed = off;
		}
	}
	if (j->mode == 0) fsm_eager_output_set_cb(j->fsm, NULL, NULL);
	return NULL;
}

static double last_walk_seconds;

int
refh_exec_eager_batch(void *vfsm, const uint8_t *base, const uint64_t *offsets, size_t n,
	int mode, int nthreads, struct fsm_b200_result *out, uint64_t *masks, size_t words,
	const uint32_t *id_of_bit, size_t nbits)
{
	struct fsm *fsm = vfsm;
	pthread_t tids[256];
	struct ejob jobs[256];
	int t, made = 0, rc = 0;
	fsm_state_t st;

	if (!fsm_all(fsm, fsm_isdfa) || !fsm_getstart(fsm, &st)) { errno = EINVAL; return -1; }
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 256) nthreads = 256;
	for (t = 0; t < nthreads; t++) {
		jobs[t].fsm = t == 0 ? fsm : fsm_clone(fsm);
		if (jobs[t].fsm == NULL) { rc = -1; break; }
		made = t + 1;
		jobs[t].base = base; jobs[t].offsets = offsets;
		jobs[t].lo = n * (size_t) t / (size_t) nthreads;
		jobs[t].hi = n * (size_t) (t + 1) / (size_t) nthreads;
		jobs[t].mode = mode; jobs[t].out = out;
		jobs[t].masks = masks; jobs[t].words = words; jobs[t].id_of_bit = id_of_bit; jobs[t].nbits = nbits;
	}
	if (rc == 0) {
		struct timespec t0, t1;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		if (nthreads == 1) {
			eworker(&jobs[0]);
		} else {
			int started = 0;
			for (t = 0; t < nthreads; t++) {
				if (pthread_create(&tids[t], NULL, eworker, &jobs[t]) != 0) { rc = -1; errno = EAGAIN; break; }
				started = t + 1;
			}
			for (t = 0; t < started; t++) pthread_join(tids[t], NULL);
		}
		clock_gettime(CLOCK_MONOTONIC, &t1);
		last_walk_seconds = (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
	}
	for (t = 1; t < made; t++) fsm_free(jobs[t].fsm);
	return rc;
}

/* Seconds the threads of the last refh_exec_eager_batch call spent walking (thread start to join):
 * the per-thread fsm_clone the harness needs -- the callback slot lives in the fsm -- is set-up a
 * multi-threaded user of the reference would pay once, not per batch, so benchmarks time this. */
double
refh_last_walk_seconds(void)
{
	return last_walk_seconds;
}

/* examples/utf8dfa/main.c restated over the same API calls (its output languages are dot / api / c
 * only, so the automaton is rebuilt here rather than parsed): one chain of fsm_addedge_literal per
 * code point from the start state (main.c:64-96), surrogates skipped (:218-224), then
 * fsm_determinise + fsm_minimise (:229-237).  [lo, hi] = 0..0x10FFFF gives the 9-state DFA that
 * accepts exactly ONE code point. */
static int
utf8_encode(int cp, char c[4])
{
	if (cp < 0) return 0;
	if (cp <= 0x7f) { c[0] = (char) cp; return 1; }
	if (cp <= 0x7ff) { c[0] = (char) ((cp >> 6) + 192); c[1] = (char) ((cp & 63) + 128); return 2; }
	if (0xd800 <= cp && cp <= 0xdfff) return 0;
	if (cp <= 0xffff) { c[0] = (char) ((cp >> 12) + 224); c[1] = (char) (((cp >> 6) & 63) + 128); c[2] = (char) ((cp & 63) + 128); return 3; }
	if (cp <= 0x10ffff) {
		c[0] = (char) ((cp >> 18) + 240); c[1] = (char) (((cp >> 12) & 63) + 128);
		c[2] = (char) (((cp >> 6) & 63) + 128); c[3] = (char) ((cp & 63) + 128);
		return 4;
	}
	return 0;
}

void *
refh_utf8dfa(int lo, int hi)
{
	struct fsm *fsm = fsm_new(NULL);
	fsm_state_t start;
	int cp;
	if (fsm == NULL) return NULL;
	if (!fsm_addstate(fsm, &start)) goto fail;
	fsm_setstart(fsm, start);
	for (cp = lo; cp <= hi; cp++) {
		char c[4];
		fsm_state_t x = start, y;
		int r, i;
		if (0xd800 <= cp && cp <= 0xdfff) continue;
		r = utf8_encode(cp, c);
		if (r == 0) goto fail;
		for (i = 0; i < r; i++) {
			if (!fsm_addstate(fsm, &y)) goto fail;
			if (!fsm_addedge_literal(fsm, x, y, c[i])) goto fail;
			x = y;
		}
		fsm_setend(fsm, x, 1);
	}
	if (!fsm_determinise(fsm)) goto fail;
	if (!fsm_minimise(fsm)) goto fail;
	return fsm;
fail:
	fsm_free(fsm);
	return NULL;
}

/* Kleene star through the reference API: an epsilon edge from every
 * end state back to the start state, and the start state becomes an end state.  The caller
 * determinises + minimises. */
int
refh_star(void *vfsm)
{
	struct fsm *fsm = vfsm;
	fsm_state_t start, s;
	if (!fsm_getstart(fsm, &start)) return 0;
	for (s = 0; s < fsm->statecount; s++) {
		if (fsm_isend(fsm, s) && !fsm_addedge_epsilon(fsm, s, start)) return 0;
	}
	fsm_setend(fsm, start, 1);
	return 1;
}

/* ---- the DFAVM bytecode engine (src/libfsm/vm.c, vm/v1.c): secondary CPU baseline + loader fixture ---- */

enum dfavm_io_result { DFAVM_IO_OK_ = 0 };
int fsm_dfavm_save(FILE *f, const struct fsm_dfavm *vm);        /* src/libfsm/vm.c:39-49 (internal, not in libfsm.syms) */

/* fsm_vm_compile + fsm_dfavm_save: the "DFAVM$" file image of a DFA (malloc'd). */
int
refh_dfavm_bytes(const void *fsm, uint8_t **out, size_t *len)
{
	struct fsm_dfavm *vm = fsm_vm_compile(fsm);
	char *buf = NULL;
	size_t n = 0;
	FILE *f;
	int rc;
	if (vm == NULL) return -1;
	f = open_memstream(&buf, &n);
	if (f == NULL) { fsm_vm_free(vm); return -1; }
	rc = fsm_dfavm_save(f, vm);
	fclose(f);
	fsm_vm_free(vm);
	if (rc != 0) { free(buf); return -1; }
	*out = (uint8_t *) buf; *len = n;
	return 0;
}

struct vjob { const struct fsm_dfavm *vm; const uint8_t *base; const uint64_t *offsets; size_t lo, hi; uint8_t *out; };

static void *
vworker(void *opaque)
{
	struct vjob *j = opaque;
	size_t i;
	for (i = j->lo; i < j->hi; i++) {
		j->out[i] = (uint8_t) (fsm_vm_match_buffer(j->vm, (const char *) j->base + j->offsets[i],
		    (size_t) (j->offsets[i + 1] - j->offsets[i])) != 0);
	}
	return NULL;
}

/* n fsm_vm_match_buffer calls (vm.c:218-229 -> vm_match_v1) over nthreads pthreads; out[i] = matched.
 * The VM is compiled once (the reference's retest does the same, runner.c:430-438). */
int
refh_vm_match_batch(const void *fsm, const uint8_t *base, const uint64_t *offsets, size_t n, int nthreads, uint8_t *out)
{
	struct fsm_dfavm *vm = fsm_vm_compile(fsm);
	pthread_t tids[256];
	struct vjob jobs[256];
	int t, started = 0, rc = 0;
	if (vm == NULL) return -1;
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 256) nthreads = 256;
	for (t = 0; t < nthreads; t++) {
		jobs[t].vm = vm; jobs[t].base = base; jobs[t].offsets = offsets; jobs[t].out = out;
		jobs[t].lo = n * (size_t) t / (size_t) nthreads;
		jobs[t].hi = n * (size_t) (t + 1) / (size_t) nthreads;
	}
	if (nthreads == 1) {
		vworker(&jobs[0]);
	} else {
		for (t = 0; t < nthreads; t++) {
			if (pthread_create(&tids[t], NULL, vworker, &jobs[t]) != 0) { rc = -1; errno = EAGAIN; break; }
			started = t + 1;
		}
		for (t = 0; t < started; t++) pthread_join(tids[t], NULL);
	}
	fsm_vm_free(vm);
	return rc;
}

