/*
 * fsm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See fsm_oracle.h.
 *
 * A from-scratch CPU restatement of the katef/libfsm hot path over the flat
 * `struct fsm_b200_desc`.  No reference source is included or copied; each function cites
 * the reference file:line whose observable behaviour it restates.  Parity PINNED against
 * the compiled reference (oracle/_ref/libfsm_ref.so) by tests/test_oracle_vs_reference.py
 * and against tests/golden/ fixtures generated from the reference.
 */
#define _POSIX_C_SOURCE 200809L
#include <errno.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "fsm_oracle.h"

#define NO_EDGE UINT32_MAX

static inline int
sym_get(const uint64_t *symbols, unsigned c)
{
	return (int) ((symbols[c >> 6] >> (c & 63)) & 1u);
}

/* ---------------------------------------------------------------------------------
 * fsm_all(fsm, fsm_isdfa) followed by fsm_getstart, as fsm_exec does before touching
 * the input (src/libfsm/exec.c:106-114).
 *  - pred/isdfa.c:36-38: no start state => not a DFA (for every state, so also for an
 *    fsm with states but no start; with zero states fsm_all is vacuously true and the
 *    following fsm_getstart fails: same verdict)
 *  - pred/isdfa.c:43-45: any epsilon edge => not a DFA
 *  - src/adt/edgeset.c:514-562: a symbol present in two groups of one state => not a DFA
 * --------------------------------------------------------------------------------- */
int
oracle_isdfa(const struct fsm_b200_desc *d)
{
	uint32_t s;

	if (!d->hasstart || d->start >= d->nstates) {
		return 0;
	}
	for (s = 0; s < d->nstates; s++) {
		uint64_t seen[4] = { 0, 0, 0, 0 };
		uint64_t g;

		if (d->eps_off != NULL && d->eps_off[s + 1] != d->eps_off[s]) {
			return 0;
		}
		for (g = d->group_off[s]; g < d->group_off[s + 1]; g++) {
			const uint64_t *sym = &d->group_symbols[4 * g];
			int w;
			for (w = 0; w < 4; w++) {
				if (seen[w] & sym[w]) {
					return 0;
				}
				seen[w] |= sym[w];
			}
		}
	}
	return 1;
}

/* edge_set_find (src/adt/edgeset.c:394-418): linear scan of the state's groups in stored
 * order, first group whose mask contains the symbol wins. */
static inline uint32_t
group_scan(const struct fsm_b200_desc *d, uint32_t state, unsigned c)
{
	uint64_t g;
	for (g = d->group_off[state]; g < d->group_off[state + 1]; g++) {
		if (sym_get(&d->group_symbols[4 * g], c)) {
			return d->group_to[g];
		}
	}
	return NO_EDGE;
}

/* eager-output CSR of a description (both NULL when it has none) */
static void
desc_eager(const struct fsm_b200_desc *d, const uint64_t **off, const uint32_t **ids)
{
	*off = NULL; *ids = NULL;
	if (d->reserved & FSM_B200_DESC_EAGER) {
		const struct fsm_b200_desc_ext *x = (const struct fsm_b200_desc_ext *) d;
		*off = x->eager_off; *ids = x->eager_ids;
	}
}

/* fsm_exec (src/libfsm/exec.c:85-167) without captures / eager outputs (neither is
 * produced by re_comp). */
int
oracle_exec(const struct fsm_b200_desc *d, const uint8_t *buf, uint64_t len,
	int validate, struct fsm_b200_result *out)
{
	uint32_t state;
	uint64_t offset = 0;

	if (validate && !oracle_isdfa(d)) {       /* exec.c:106-109 */
		errno = EINVAL;
		return -1;
	}
	if (!d->hasstart) {                       /* exec.c:111-114 */
		errno = EINVAL;
		return -1;
	}
	state = d->start;

	while (offset < len) {                    /* exec.c:132 (getc != EOF) */
		uint32_t next = group_scan(d, state, buf[offset]);
		if (next == NO_EDGE) {                /* exec.c:133-138: stop, input not drained */
			out->ret = 0;
			out->end = state;
			out->consumed = offset;
			return 0;
		}
		state = next;
		offset++;                             /* exec.c:150 */
	}

	out->end = state;
	out->consumed = offset;
	out->ret = d->is_end[state] ? 1 : 0;      /* exec.c:153-166 */
	return out->ret;
}

static void
fire(uint32_t id, uint32_t *fired, size_t cap, size_t *n)
{
	size_t i, pos = *n < cap ? *n : cap;
	for (i = 0; i < pos; i++) if (fired[i] == id) return;
	/* keep the stored prefix sorted: insertion */
	if (*n < cap) {
		size_t j = *n;
		while (j > 0 && fired[j - 1] > id) { fired[j] = fired[j - 1]; j--; }
		fired[j] = id;
	}
	(*n)++;
}

int
oracle_exec_eager(const struct fsm_b200_desc *d, const uint8_t *buf, uint64_t len,
	struct fsm_b200_result *out, uint32_t *fired, size_t cap, size_t *nfired)
{
	const uint64_t *xoff; const uint32_t *xids;
	uint32_t state;
	uint64_t offset = 0, q;

	*nfired = 0;
	desc_eager(d, &xoff, &xids);
	if (!oracle_isdfa(d) || !d->hasstart) {   /* exec.c:106-114 */
		errno = EINVAL;
		return -1;
	}
	state = d->start;
	if (xoff != NULL) {                       /* exec.c:126-130: the start state fires too */
		for (q = xoff[state]; q < xoff[state + 1]; q++) fire(xids[q], fired, cap, nfired);
	}
	while (offset < len) {
		uint32_t next = group_scan(d, state, buf[offset]);
		if (next == NO_EDGE) {
			out->ret = 0; out->end = state; out->consumed = offset;
			return 0;
		}
		state = next;
		if (xoff != NULL) {                   /* exec.c:140-144 */
			for (q = xoff[state]; q < xoff[state + 1]; q++) fire(xids[q], fired, cap, nfired);
		}
		offset++;
	}
	out->end = state;
	out->consumed = offset;
	out->ret = d->is_end[state] ? 1 : 0;
	return out->ret;
}

struct batch_job {
	const struct fsm_b200_desc *d;
	const uint8_t *base;
	const uint64_t *offsets;
	size_t lo, hi;
	int validate_each;
	struct fsm_b200_result *out;
};

static void *
batch_worker(void *opaque)
{
	struct batch_job *j = opaque;
	size_t i;
	for (i = j->lo; i < j->hi; i++) {
		(void) oracle_exec(j->d, j->base + j->offsets[i],
		    j->offsets[i + 1] - j->offsets[i], j->validate_each, &j->out[i]);
	}
	return NULL;
}

int
oracle_exec_batch(const struct fsm_b200_desc *d, const uint8_t *base,
	const uint64_t *offsets, size_t n, int validate_each, int nthreads,
	struct fsm_b200_result *out)
{
	pthread_t *tids;
	struct batch_job *jobs;
	int t, started = 0;

	if (!oracle_isdfa(d)) {
		errno = EINVAL;
		return -1;
	}
	if (nthreads < 1) {
		nthreads = 1;
	}
	if ((size_t) nthreads > n && n > 0) {
		nthreads = (int) n;
	}
	tids = malloc(sizeof *tids * (size_t) nthreads);
	jobs = malloc(sizeof *jobs * (size_t) nthreads);
	if (tids == NULL || jobs == NULL) {
		free(tids); free(jobs);
		errno = ENOMEM;
		return -1;
	}
	for (t = 0; t < nthreads; t++) {
		jobs[t].d = d; jobs[t].base = base; jobs[t].offsets = offsets;
		jobs[t].lo = n * (size_t) t / (size_t) nthreads;
		jobs[t].hi = n * (size_t) (t + 1) / (size_t) nthreads;
		jobs[t].validate_each = validate_each;
		jobs[t].out = out;
		if (nthreads == 1) {
			batch_worker(&jobs[t]);
		} else if (pthread_create(&tids[t], NULL, batch_worker, &jobs[t]) != 0) {
			batch_worker(&jobs[t]);
			tids[t] = (pthread_t) 0;
			continue;
		} else {
			started++;
		}
	}
	if (nthreads > 1) {
		for (t = 0; t < nthreads; t++) {
			if (tids[t] != (pthread_t) 0) {
				pthread_join(tids[t], NULL);
			}
		}
	}
	(void) started;
	free(tids); free(jobs);
	return 0;
}

void
oracle_flatten(const struct fsm_b200_desc *d, uint32_t *table)
{
	uint32_t s;
	unsigned c;
	for (s = 0; s < d->nstates; s++) {
		for (c = 0; c < 256; c++) {
			table[(size_t) s * 256 + c] = group_scan(d, s, c);
		}
	}
}

/* ---------------------------------------------------------------------------------
 * Epsilon closure (src/libfsm/closure.c:130-190, :24-128): for every state the set of
 * states reachable by zero or more epsilon edges, itself included, ascending.
 * --------------------------------------------------------------------------------- */
static int
cmp_u32(const void *a, const void *b)
{
	uint32_t x = *(const uint32_t *) a, y = *(const uint32_t *) b;
	return (x > y) - (x < y);
}

int
oracle_epsilon_closure(const struct fsm_b200_desc *d, uint64_t **off_out, uint32_t **to_out)
{
	const uint32_t n = d->nstates;
	uint64_t *off = malloc(sizeof *off * ((size_t) n + 1));
	uint32_t *to = NULL, *stack = malloc(sizeof *stack * ((size_t) n + 1));
	uint32_t *mark = calloc((size_t) n + 1, sizeof *mark);   /* generation stamps */
	size_t used = 0, cap = 0;
	uint32_t s;

	if (off == NULL || stack == NULL || mark == NULL) {
		goto oom;
	}
	for (s = 0; s < n; s++) {
		size_t sp = 0, begin = used;
		off[s] = used;
		stack[sp++] = s;
		mark[s] = s + 1;
		while (sp > 0) {
			uint32_t u = stack[--sp];
			uint64_t e;
			if (used == cap) {
				size_t ncap = cap ? cap * 2 : 1024;
				uint32_t *nt = realloc(to, sizeof *nt * ncap);
				if (nt == NULL) goto oom;
				to = nt; cap = ncap;
			}
			to[used++] = u;
			if (d->eps_off == NULL) continue;
			for (e = d->eps_off[u]; e < d->eps_off[u + 1]; e++) {
				uint32_t v = d->eps_to[e];
				if (mark[v] != s + 1) {
					mark[v] = s + 1;
					stack[sp++] = v;
				}
			}
		}
		qsort(to + begin, used - begin, sizeof *to, cmp_u32);
	}
	off[n] = used;
	free(stack); free(mark);
	if (to == NULL) {
		to = malloc(sizeof *to);
		if (to == NULL) { free(off); errno = ENOMEM; return -1; }
	}
	*off_out = off; *to_out = to;
	return 0;
oom:
	free(off); free(to); free(stack); free(mark);
	errno = ENOMEM;
	return -1;
}

/* ---------------------------------------------------------------------------------
 * Subset construction.
 *
 * The reference first folds epsilon closures into the labelled edges
 * (fsm_remove_epsilons, src/libfsm/epsilons.c:180-254): state s gets the union of the
 * edge groups, the end bit and the end ids of every state in closure(s).  It then runs
 * the subset construction on that epsilon-free NFA starting from the set {start}
 * (determinise.c:88-104): a DFA state is a set of NFA states that is NOT re-closed, so
 * two different sets with equal closures stay distinct states.  This restatement keeps
 * exactly that formulation so that state counts agree with the reference; only the
 * numbering differs (BFS over symbols here; LIFO worklist + analysis order there).
 * --------------------------------------------------------------------------------- */
struct vec32 { uint32_t *a; size_t n, cap; };

static int
vec32_push(struct vec32 *v, uint32_t x)
{
	if (v->n == v->cap) {
		size_t ncap = v->cap ? v->cap * 2 : 16;
		uint32_t *na = realloc(v->a, sizeof *na * ncap);
		if (na == NULL) return 0;
		v->a = na; v->cap = ncap;
	}
	v->a[v->n++] = x;
	return 1;
}

struct vec64 { uint64_t *a; size_t n, cap; };

static int
vec64_push(struct vec64 *v, uint64_t x)
{
	if (v->n == v->cap) {
		size_t ncap = v->cap ? v->cap * 2 : 16;
		uint64_t *na = realloc(v->a, sizeof *na * ncap);
		if (na == NULL) return 0;
		v->a = na; v->cap = ncap;
	}
	v->a[v->n++] = x;
	return 1;
}

/* interning pool of sorted u32 sets */
struct pool {
	struct vec32 buf;        /* concatenated sets */
	struct vec64 off;        /* set i = buf[off[i] .. off[i+1]) */
	uint32_t *htab;          /* open addressing, value = set id + 1 */
	size_t hcap;             /* power of two */
};

static uint64_t
hash_set(const uint32_t *a, size_t n)
{
	uint64_t h = 0xcbf29ce484222325ull ^ (uint64_t) n;
	size_t i;
	for (i = 0; i < n; i++) {
		h ^= a[i];
		h *= 0x100000001b3ull;
		h ^= h >> 29;
	}
	return h;
}

static int
pool_grow(struct pool *p)
{
	size_t ncap = p->hcap ? p->hcap * 2 : 1024, i;
	uint32_t *nt = calloc(ncap, sizeof *nt);
	if (nt == NULL) return 0;
	for (i = 0; i + 1 < p->off.n; i++) {
		const uint32_t *a = p->buf.a + p->off.a[i];
		size_t n = (size_t) (p->off.a[i + 1] - p->off.a[i]);
		size_t h = (size_t) hash_set(a, n) & (ncap - 1);
		while (nt[h] != 0) h = (h + 1) & (ncap - 1);
		nt[h] = (uint32_t) i + 1;
	}
	free(p->htab);
	p->htab = nt; p->hcap = ncap;
	return 1;
}

/* returns id, sets *isnew; (uint32_t)-1 on OOM */
static uint32_t
pool_intern(struct pool *p, const uint32_t *a, size_t n, int *isnew)
{
	size_t h, count = p->off.n ? p->off.n - 1 : 0, i;
	if (p->off.n == 0 && !vec64_push(&p->off, 0)) return (uint32_t) -1;
	if ((count + 1) * 2 > p->hcap && !pool_grow(p)) return (uint32_t) -1;
	h = (size_t) hash_set(a, n) & (p->hcap - 1);
	while (p->htab[h] != 0) {
		uint32_t id = p->htab[h] - 1;
		size_t m = (size_t) (p->off.a[id + 1] - p->off.a[id]);
		if (m == n && memcmp(p->buf.a + p->off.a[id], a, n * sizeof *a) == 0) {
			*isnew = 0;
			return id;
		}
		h = (h + 1) & (p->hcap - 1);
	}
	for (i = 0; i < n; i++) {
		if (!vec32_push(&p->buf, a[i])) return (uint32_t) -1;
	}
	if (!vec64_push(&p->off, p->buf.n)) return (uint32_t) -1;
	p->htab[h] = (uint32_t) count + 1;
	*isnew = 1;
	return (uint32_t) count;
}


static size_t
sort_unique(uint32_t *a, size_t n)
{
	size_t i, w = 0;
	if (n < 2) return n;
	qsort(a, n, sizeof *a, cmp_u32);
	for (i = 0; i < n; i++) {
		if (w == 0 || a[w - 1] != a[i]) a[w++] = a[i];
	}
	return w;
}

int
oracle_determinise(const struct fsm_b200_desc *nfa, size_t state_limit,
	struct oracle_owned_desc *out)
{
	const uint32_t n = nfa->nstates;
	uint64_t *cl_off = NULL; uint32_t *cl_to = NULL;
	uint64_t *adj_off = NULL;          /* per NFA state: expanded (symbol,to) edges */
	uint8_t *adj_sym = NULL; uint32_t *adj_to = NULL;
	uint8_t *aend = NULL;              /* augmented end bit */
	struct vec32 symlist[256];
	struct pool pool;
	struct vec64 o_goff = {0}, o_gsym = {0}, o_eoff = {0};
	struct vec32 o_gto = {0}, o_eid = {0}, o_end = {0}, tmp = {0};
	struct vec64 o_xoff = {0};
	struct vec32 o_xid = {0}, tmp2 = {0};
	const uint64_t *xoff; const uint32_t *xids;
	int rc = -1;
	uint32_t s;
	size_t c;
	uint8_t *reach = NULL;

	memset(out, 0, sizeof *out);
	memset(symlist, 0, sizeof symlist);
	memset(&pool, 0, sizeof pool);

	desc_eager(nfa, &xoff, &xids);
	if (oracle_epsilon_closure(nfa, &cl_off, &cl_to) != 0) return -1;

	/* determinise.c:65-68 */
	if (state_limit != 0 && n > state_limit) { rc = 1; goto done; }

	/* mark_states_reachable_by_label (epsilons.c:296-323) */
	reach = calloc((size_t) n + 1, 1);
	aend = calloc((size_t) n + 1, 1);
	adj_off = calloc((size_t) n + 2, sizeof *adj_off);
	if (!reach || !aend || !adj_off) goto oom;
	for (s = 0; s < n; s++) {
		uint64_t g;
		for (g = nfa->group_off[s]; g < nfa->group_off[s + 1]; g++) {
			reach[nfa->group_to[g]] = 1;
		}
	}
	if (nfa->hasstart) reach[nfa->start] = 1;

	/* augmented, symbol-expanded adjacency: two passes (count, fill) */
	for (int pass = 0; pass < 2; pass++) {
		uint64_t total = 0;
		for (s = 0; s < n; s++) {
			uint64_t k;
			if (pass == 1) total = adj_off[s];
			else adj_off[s] = total;
			if (!reach[s]) continue;
			for (k = cl_off[s]; k < cl_off[s + 1]; k++) {
				uint32_t es = cl_to[k];
				uint64_t g;
				if (nfa->is_end[es]) aend[s] = 1;
				for (g = nfa->group_off[es]; g < nfa->group_off[es + 1]; g++) {
					const uint64_t *sym = &nfa->group_symbols[4 * g];
					for (c = 0; c < 256; c++) {
						if (!sym_get(sym, (unsigned) c)) continue;
						if (pass == 1) {
							adj_sym[total] = (uint8_t) c;
							adj_to[total] = nfa->group_to[g];
						}
						total++;
					}
				}
			}
		}
		if (pass == 0) {
			adj_off[n] = total;
			adj_sym = malloc(total + 1);
			adj_to = malloc(sizeof *adj_to * (total + 1));
			if (!adj_sym || !adj_to) goto oom;
		}
	}
	/* states not reachable by label keep their own end bit (they are garbage) */
	for (s = 0; s < n; s++) if (!reach[s]) aend[s] = nfa->is_end[s];

	if (!nfa->hasstart) {   /* determinise.c:88-91: nothing to do, fsm left as is */
		rc = 0;
		out->desc.nstates = 0;
		goto done;
	}

	{
		int isnew;
		uint32_t st = nfa->start;
		if (pool_intern(&pool, &st, 1, &isnew) == (uint32_t) -1) goto oom;
	}
	if (!vec64_push(&o_goff, 0) || !vec64_push(&o_eoff, 0) || !vec64_push(&o_xoff, 0)) goto oom;

	/* BFS: DFA state id == interned set id (discovery order) */
	for (uint32_t cur = 0; cur + 1 < pool.off.n; cur++) {
		const size_t mb = (size_t) pool.off.a[cur], me = (size_t) pool.off.a[cur + 1];
		uint32_t dst_of_sym[256];
		uint8_t used[256];
		int is_end = 0;
		size_t i;

		memset(used, 0, sizeof used);
		for (c = 0; c < 256; c++) symlist[c].n = 0;
		tmp.n = 0;
		tmp2.n = 0;

		for (i = mb; i < me; i++) {
			uint32_t m = pool.buf.a[i];
			uint64_t e, k;
			if (xoff != NULL) {
				/* eager outputs: every state of the member's closure (epsilons.c:221-253) */
				for (k = cl_off[m]; k < cl_off[m + 1]; k++) {
					uint64_t q;
					for (q = xoff[cl_to[k]]; q < xoff[cl_to[k] + 1]; q++) {
						if (!vec32_push(&tmp2, xids[q])) goto oom;
					}
				}
			}
			if (aend[m]) {
				is_end = 1;
				/* end ids: union over closure members that are end states
				 * (epsilons.c:529-569 then endids.c:782-826) */
				for (k = cl_off[m]; k < cl_off[m + 1]; k++) {
					uint32_t es = cl_to[k];
					uint64_t q;
					if (!nfa->is_end[es] || nfa->endid_off == NULL) continue;
					for (q = nfa->endid_off[es]; q < nfa->endid_off[es + 1]; q++) {
						if (!vec32_push(&tmp, nfa->endids[q])) goto oom;
					}
				}
			}
			for (e = adj_off[m]; e < adj_off[m + 1]; e++) {
				if (!vec32_push(&symlist[adj_sym[e]], adj_to[e])) goto oom;
				used[adj_sym[e]] = 1;
			}
		}
		if (!vec32_push(&o_end, (uint32_t) is_end)) goto oom;
		tmp.n = sort_unique(tmp.a, tmp.n);
		for (i = 0; i < tmp.n; i++) if (!vec32_push(&o_eid, tmp.a[i])) goto oom;
		if (!vec64_push(&o_eoff, o_eid.n)) goto oom;
		tmp2.n = sort_unique(tmp2.a, tmp2.n);
		for (i = 0; i < tmp2.n; i++) if (!vec32_push(&o_xid, tmp2.a[i])) goto oom;
		if (!vec64_push(&o_xoff, o_xid.n)) goto oom;

		for (c = 0; c < 256; c++) {
			int isnew;
			dst_of_sym[c] = NO_EDGE;
			if (!used[c]) continue;
			symlist[c].n = sort_unique(symlist[c].a, symlist[c].n);
			dst_of_sym[c] = pool_intern(&pool, symlist[c].a, symlist[c].n, &isnew);
			if (dst_of_sym[c] == (uint32_t) -1) goto oom;
			if (isnew && state_limit != 0 && (pool.off.n - 2) > state_limit) {
				/* determinise.c:166-169: checked with the count BEFORE adding */
				rc = 1;
				goto done;
			}
		}
		/* groups: one per distinct destination, ascending destination
		 * (edge_set keeps groups sorted by .to, src/adt/edgeset.c:283-373) */
		{
			uint32_t dsts[256]; size_t nd = 0, j;
			for (c = 0; c < 256; c++) if (dst_of_sym[c] != NO_EDGE) dsts[nd++] = dst_of_sym[c];
			nd = sort_unique(dsts, nd);
			for (j = 0; j < nd; j++) {
				uint64_t sym[4] = {0, 0, 0, 0};
				for (c = 0; c < 256; c++) {
					if (dst_of_sym[c] == dsts[j]) sym[c >> 6] |= 1ull << (c & 63);
				}
				if (!vec64_push(&o_gsym, sym[0]) || !vec64_push(&o_gsym, sym[1]) ||
				    !vec64_push(&o_gsym, sym[2]) || !vec64_push(&o_gsym, sym[3]) ||
				    !vec32_push(&o_gto, dsts[j])) goto oom;
			}
			if (!vec64_push(&o_goff, o_gto.n)) goto oom;
		}
	}

	{
		const uint32_t nd = (uint32_t) (pool.off.n - 1);
		uint8_t *is_end = malloc((size_t) nd + 1);
		uint32_t i;
		if (!is_end) goto oom;
		for (i = 0; i < nd; i++) is_end[i] = (uint8_t) o_end.a[i];
		out->desc.nstates = nd;
		out->desc.start = 0;
		out->desc.hasstart = 1;
		out->desc.is_end = is_end;
		out->desc.group_off = o_goff.a;
		out->desc.group_symbols = o_gsym.a ? o_gsym.a : calloc(4, sizeof(uint64_t));
		out->desc.group_to = o_gto.a ? o_gto.a : calloc(1, sizeof(uint32_t));
		out->desc.eps_off = NULL;
		out->desc.eps_to = NULL;
		out->desc.endid_off = o_eoff.a;
		out->desc.endids = o_eid.a ? o_eid.a : calloc(1, sizeof(uint32_t));
		out->blocks[0] = is_end;
		out->blocks[1] = (void *) out->desc.group_off;
		out->blocks[2] = (void *) out->desc.group_symbols;
		out->blocks[3] = (void *) out->desc.group_to;
		out->blocks[4] = (void *) out->desc.endid_off;
		out->blocks[5] = (void *) out->desc.endids;
		o_goff.a = NULL; o_gsym.a = NULL; o_gto.a = NULL; o_eoff.a = NULL; o_eid.a = NULL;
		if (xoff != NULL && o_xid.n > 0) {
			out->eager_off = o_xoff.a; out->eager_ids = o_xid.a;
			out->blocks[6] = o_xoff.a; out->blocks[7] = o_xid.a;
			o_xoff.a = NULL; o_xid.a = NULL;
		}
		rc = 0;
	}
	goto done;
oom:
	errno = ENOMEM;
	rc = -1;
done:
	free(o_xoff.a); free(o_xid.a); free(tmp2.a);
	free(cl_off); free(cl_to); free(adj_off); free(adj_sym); free(adj_to);
	free(aend); free(reach);
	for (c = 0; c < 256; c++) free(symlist[c].a);
	free(pool.buf.a); free(pool.off.a); free(pool.htab);
	free(o_goff.a); free(o_gsym.a); free(o_gto.a); free(o_eoff.a); free(o_eid.a);
	free(o_end.a); free(tmp.a);
	return rc;
}

void
oracle_desc_free(struct oracle_owned_desc *d)
{
	size_t i;
	for (i = 0; i < sizeof d->blocks / sizeof d->blocks[0]; i++) {
		free(d->blocks[i]);
		d->blocks[i] = NULL;
	}
	memset(&d->desc, 0, sizeof d->desc);
	d->eager_off = NULL; d->eager_ids = NULL;
}

/* ---------------------------------------------------------------------------------
 * Minimisation (Moore refinement on the trimmed DFA).
 * --------------------------------------------------------------------------------- */
struct sigrow { uint32_t state; const uint32_t *sig; size_t len; };

enum { SIGLEN = 257 };      /* class of the state + class of each of its 256 successors */

static int
cmp_sigrow(const void *a, const void *b)
{
	const struct sigrow *x = a, *y = b;
	int c = memcmp(x->sig, y->sig, SIGLEN * sizeof(uint32_t));
	if (c != 0) return c;
	return (x->state > y->state) - (x->state < y->state);
}

static int minimise_impl(const struct fsm_b200_desc *d, const uint32_t *given_cls0, struct oracle_owned_desc *out);

int
oracle_minimise(const struct fsm_b200_desc *d, struct oracle_owned_desc *out)
{
	return minimise_impl(d, NULL, out);
}

/* Test hook: the same refinement started from initial classes computed elsewhere (one id per
 * input state, 0 = plain) -- lets tests check the PRODUCT's host code for the initial partition
 * (libfsm_b200/csrc/eager_host.h, compiled for the CPU by oracle/eager_host_test.cpp). */
int
oracle_minimise_from_classes(const struct fsm_b200_desc *d, const uint32_t *cls0, struct oracle_owned_desc *out)
{
	return minimise_impl(d, cls0, out);
}

static int
minimise_impl(const struct fsm_b200_desc *d, const uint32_t *given_cls0, struct oracle_owned_desc *out)
{
	const uint32_t n = d->nstates;
	uint32_t *table = NULL, *newid = NULL, *cls = NULL, *ncls_arr = NULL, *sig = NULL, *rep = NULL;
	uint8_t *reach = NULL, *co = NULL;
	struct sigrow *rows = NULL;
	struct vec64 o_goff = {0}, o_gsym = {0}, o_eoff = {0}, o_xoff = {0};
	struct vec32 o_gto = {0}, o_eid = {0}, o_xid = {0}, xtmp = {0};
	const uint64_t *xoff; const uint32_t *xids;
	uint32_t *dist = NULL;
	uint8_t *collected = NULL, *broken = NULL;
	uint32_t m = 0, ncls = 0, s;
	unsigned c;
	int rc = -1, changed;

	memset(out, 0, sizeof *out);
	desc_eager(d, &xoff, &xids);
	if (!oracle_isdfa(d)) { errno = EINVAL; return -1; }
	table = malloc(sizeof *table * (size_t) n * 256 + 4);
	reach = calloc((size_t) n + 1, 1); co = calloc((size_t) n + 1, 1);
	newid = malloc(sizeof *newid * ((size_t) n + 1));
	if (!table || !reach || !co || !newid) goto oom;
	oracle_flatten(d, table);

	/* trim: reachable from start ... */
	reach[d->start] = 1;
	do {
		changed = 0;
		for (s = 0; s < n; s++) {
			if (!reach[s]) continue;
			for (c = 0; c < 256; c++) {
				uint32_t t = table[(size_t) s * 256 + c];
				if (t != NO_EDGE && !reach[t]) { reach[t] = 1; changed = 1; }
			}
		}
	} while (changed);
	/* ... and able to reach an end state */
	for (s = 0; s < n; s++) co[s] = d->is_end[s] ? 1 : 0;
	do {
		changed = 0;
		for (s = 0; s < n; s++) {
			if (co[s]) continue;
			for (c = 0; c < 256; c++) {
				uint32_t t = table[(size_t) s * 256 + c];
				if (t != NO_EDGE && co[t]) { co[s] = 1; changed = 1; break; }
			}
		}
	} while (changed);
	for (s = 0; s < n; s++) newid[s] = (reach[s] && co[s]) ? m++ : NO_EDGE;
	if (m == 0 || newid[d->start] == NO_EDGE) {
		/* nothing left (minimise.c:98-101): an empty fsm */
		uint64_t *z = calloc(2, sizeof *z), *z2 = calloc(2, sizeof *z2);
		if (!z || !z2) { free(z); free(z2); goto oom; }
		out->desc.nstates = 0; out->desc.hasstart = 0;
		out->desc.group_off = z; out->desc.endid_off = z2;
		out->blocks[0] = z; out->blocks[1] = z2;
		rc = 0;
		goto done;
	}

	/* Moore refinement over the kept states; signature = (class, class of every successor) */
	cls = malloc(sizeof *cls * m); ncls_arr = malloc(sizeof *ncls_arr * m);
	sig = malloc(sizeof *sig * (size_t) m * 257); rows = malloc(sizeof *rows * m);
	if (!cls || !ncls_arr || !sig || !rows) goto oom;
	{
		/* initial classes: non-end states together; end states by end-id set */
		uint32_t *orig = malloc(sizeof *orig * m);
		if (!orig) goto oom;
		for (s = 0; s < n; s++) if (newid[s] != NO_EDGE) orig[newid[s]] = s;
		if (xoff != NULL) {
			/* Eager outputs split classes too (same_end_metadata, minimise.c:705-731) -- but the
			 * reference only LOOKS at a state's eager ids while it walks its initial class (states of
			 * equal shortest distance to an end state, listed in descending state order) and stops at
			 * the first state that is neither an end state nor has eager outputs
			 * (minimise.c:771-782): ids on states behind it are not seen, those states are merged as
			 * if they had none, and fsm_consolidate later gives the merged state the union. */
			uint32_t i, round;
			dist = malloc(sizeof *dist * m); collected = calloc(m, 1); broken = calloc((size_t) m + 1, 1);
			if (!dist || !collected || !broken) { free(orig); goto oom; }
			for (i = 0; i < m; i++) dist[i] = d->is_end[orig[i]] ? 0 : NO_EDGE;
			for (round = 1; ; round++) {             /* level-synchronous backward BFS */
				int any = 0;
				for (i = 0; i < m; i++) {
					if (dist[i] != NO_EDGE) continue;
					for (c = 0; c < 256; c++) {
						uint32_t t = table[(size_t) orig[i] * 256 + c];
						if (t != NO_EDGE && newid[t] != NO_EDGE && dist[newid[t]] == round - 1) { any = 1; dist[i] = NO_EDGE - 1; break; }
					}
				}
				for (i = 0; i < m; i++) if (dist[i] == NO_EDGE - 1) dist[i] = round;
				if (!any) break;
			}
			for (i = m; i-- > 0; ) {
				const uint32_t si = orig[i];
				if (dist[i] == NO_EDGE || broken[dist[i]]) continue;
				if (d->is_end[si] || xoff[si + 1] > xoff[si]) collected[i] = 1;
				else broken[dist[i]] = 1;
			}
		}
#define EAGER_LEN(i) ((xoff != NULL && collected[i]) ? (size_t) (xoff[orig[i] + 1] - xoff[orig[i]]) : 0)
		for (uint32_t i = 0; i < m; i++) {
			uint32_t si = orig[i], k;
			if (given_cls0 != NULL) { cls[i] = given_cls0[si]; continue; }
			cls[i] = NO_EDGE;
			if (!d->is_end[si] && EAGER_LEN(i) == 0) { cls[i] = 0; continue; }
			for (k = 0; k < i; k++) {
				uint32_t sk = orig[k];
				size_t li, lk;
				if (d->is_end[sk] != d->is_end[si] || cls[k] == NO_EDGE || cls[k] == 0) continue;
				li = (d->is_end[si] && d->endid_off) ? (size_t) (d->endid_off[si + 1] - d->endid_off[si]) : 0;
				lk = (d->is_end[sk] && d->endid_off) ? (size_t) (d->endid_off[sk + 1] - d->endid_off[sk]) : 0;
				if (li != lk || (li != 0 && memcmp(d->endids + d->endid_off[si], d->endids + d->endid_off[sk], li * sizeof(uint32_t)) != 0)) continue;
				if (EAGER_LEN(i) != EAGER_LEN(k)) continue;
				if (EAGER_LEN(i) != 0 && memcmp(xids + xoff[si], xids + xoff[sk], EAGER_LEN(i) * sizeof(uint32_t)) != 0) continue;
				cls[i] = cls[k];
				break;
			}
			if (cls[i] == NO_EDGE) cls[i] = 1 + i;      /* fresh id, distinct from 0 */
		}
#undef EAGER_LEN
		for (;;) {
			uint32_t cnt = 0, i;
			for (i = 0; i < m; i++) {
				uint32_t *row = sig + (size_t) i * 257;
				row[0] = cls[i];
				for (c = 0; c < 256; c++) {
					uint32_t t = table[(size_t) orig[i] * 256 + c];
					row[1 + c] = (t == NO_EDGE || newid[t] == NO_EDGE) ? NO_EDGE : cls[newid[t]];
				}
				rows[i].state = i; rows[i].sig = row; rows[i].len = 257;
			}
			qsort(rows, m, sizeof *rows, cmp_sigrow);
			/* class id = smallest member (rows are sorted by signature then state) */
			for (i = 0; i < m; i++) {
				if (i == 0 || memcmp(rows[i].sig, rows[i - 1].sig, 257 * sizeof(uint32_t)) != 0) {
					cnt++;
					ncls_arr[rows[i].state] = rows[i].state;
				} else {
					ncls_arr[rows[i].state] = ncls_arr[rows[i - 1].state];
				}
			}
			memcpy(cls, ncls_arr, sizeof *cls * m);
			if (cnt == ncls) break;
			ncls = cnt;
		}
		/* renumber classes densely in order of their smallest member */
		rep = malloc(sizeof *rep * m);
		if (!rep) { free(orig); goto oom; }
		{
			uint32_t next = 0;
			for (uint32_t i = 0; i < m; i++) rep[i] = NO_EDGE;
			for (uint32_t i = 0; i < m; i++) if (cls[i] == i) rep[i] = next++;
			ncls = next;
		}
		/* emit */
		{
			uint8_t *is_end = calloc((size_t) ncls + 1, 1);
			uint32_t i;
			if (!is_end || !vec64_push(&o_goff, 0) || !vec64_push(&o_eoff, 0) || !vec64_push(&o_xoff, 0)) { free(orig); free(is_end); goto oom; }
			for (i = 0; i < m; i++) {
				uint32_t dst_of_sym[256], dsts[256];
				size_t nd = 0, j;
				if (cls[i] != i) continue;
				is_end[rep[i]] = d->is_end[orig[i]];
				for (c = 0; c < 256; c++) {
					uint32_t t = table[(size_t) orig[i] * 256 + c];
					dst_of_sym[c] = (t == NO_EDGE || newid[t] == NO_EDGE) ? NO_EDGE : rep[cls[newid[t]]];
					if (dst_of_sym[c] != NO_EDGE) dsts[nd++] = dst_of_sym[c];
				}
				nd = sort_unique(dsts, nd);
				for (j = 0; j < nd; j++) {
					uint64_t sym[4] = {0, 0, 0, 0};
					for (c = 0; c < 256; c++) if (dst_of_sym[c] == dsts[j]) sym[c >> 6] |= 1ull << (c & 63);
					if (!vec64_push(&o_gsym, sym[0]) || !vec64_push(&o_gsym, sym[1]) || !vec64_push(&o_gsym, sym[2]) ||
					    !vec64_push(&o_gsym, sym[3]) || !vec32_push(&o_gto, dsts[j])) { free(orig); free(is_end); goto oom; }
				}
				if (!vec64_push(&o_goff, o_gto.n)) { free(orig); free(is_end); goto oom; }
				if (d->is_end[orig[i]] && d->endid_off) {
					uint64_t q;
					for (q = d->endid_off[orig[i]]; q < d->endid_off[orig[i] + 1]; q++) {
						if (!vec32_push(&o_eid, d->endids[q])) { free(orig); free(is_end); goto oom; }
					}
				}
				if (!vec64_push(&o_eoff, o_eid.n)) { free(orig); free(is_end); goto oom; }
				if (xoff != NULL) {
					/* fsm_consolidate (consolidate.c:306-315): the union over every merged state */
					uint32_t k; uint64_t q; size_t z;
					xtmp.n = 0;
					for (k = 0; k < m; k++) {
						if (cls[k] != i) continue;
						for (q = xoff[orig[k]]; q < xoff[orig[k] + 1]; q++) {
							if (!vec32_push(&xtmp, xids[q])) { free(orig); free(is_end); goto oom; }
						}
					}
					xtmp.n = sort_unique(xtmp.a, xtmp.n);
					for (z = 0; z < xtmp.n; z++) if (!vec32_push(&o_xid, xtmp.a[z])) { free(orig); free(is_end); goto oom; }
				}
				if (!vec64_push(&o_xoff, o_xid.n)) { free(orig); free(is_end); goto oom; }
			}
			if (xoff != NULL && o_xid.n > 0) {
				out->eager_off = o_xoff.a; out->eager_ids = o_xid.a;
				out->blocks[6] = o_xoff.a; out->blocks[7] = o_xid.a;
				o_xoff.a = NULL; o_xid.a = NULL;
			}
			out->desc.nstates = ncls;
			out->desc.start = rep[cls[newid[d->start]]];
			out->desc.hasstart = 1;
			out->desc.is_end = is_end;
			out->desc.group_off = o_goff.a;
			out->desc.group_symbols = o_gsym.a ? o_gsym.a : calloc(4, sizeof(uint64_t));
			out->desc.group_to = o_gto.a ? o_gto.a : calloc(1, sizeof(uint32_t));
			out->desc.endid_off = o_eoff.a;
			out->desc.endids = o_eid.a ? o_eid.a : calloc(1, sizeof(uint32_t));
			out->blocks[0] = is_end;
			out->blocks[1] = (void *) out->desc.group_off;
			out->blocks[2] = (void *) out->desc.group_symbols;
			out->blocks[3] = (void *) out->desc.group_to;
			out->blocks[4] = (void *) out->desc.endid_off;
			out->blocks[5] = (void *) out->desc.endids;
			o_goff.a = NULL; o_gsym.a = NULL; o_gto.a = NULL; o_eoff.a = NULL; o_eid.a = NULL;
		}
		free(orig);
		rc = 0;
	}
	goto done;
oom:
	errno = ENOMEM;
	rc = -1;
done:
	free(table); free(reach); free(co); free(newid); free(cls); free(ncls_arr); free(sig); free(rows); free(rep);
	free(o_goff.a); free(o_gsym.a); free(o_gto.a); free(o_eoff.a); free(o_eid.a);
	free(o_xoff.a); free(o_xid.a); free(xtmp.a); free(dist); free(collected); free(broken);
	return rc;
}

uint32_t
oracle_canonicalise(const struct fsm_b200_desc *d, uint32_t *canon_table,
	uint32_t *canon_of_state)
{
	const uint32_t n = d->nstates;
	uint32_t *order, ncanon = 0, head = 0, s;
	unsigned c;

	if (!oracle_isdfa(d)) return (uint32_t) -1;
	order = malloc(sizeof *order * ((size_t) n + 1));
	if (order == NULL) return (uint32_t) -1;
	for (s = 0; s < n; s++) canon_of_state[s] = NO_EDGE;
	canon_of_state[d->start] = ncanon;
	order[ncanon++] = d->start;
	while (head < ncanon) {
		uint32_t u = order[head];
		for (c = 0; c < 256; c++) {
			uint32_t v = group_scan(d, u, c);
			if (v != NO_EDGE && canon_of_state[v] == NO_EDGE) {
				canon_of_state[v] = ncanon;
				order[ncanon++] = v;
			}
			canon_table[(size_t) head * 256 + c] = (v == NO_EDGE) ? NO_EDGE : canon_of_state[v];
		}
		head++;
	}
	free(order);
	return ncanon;
}
