/*
 * refnum.h -- the reference's DFA state numbering, restated as a data-parallel computation.
 *
 * fsm_determinise numbers DFA states in the order its LIFO worklist discovers them
 * (src/libfsm/determinise.c:118-185), and the order in which one DFA state lists its
 * successors is the entry order of the "analysis result" built for its NFA-state set by
 *   - cache_single_state_analysis (determinise.c:1056-1335) for every member, and
 *   - a pairwise tournament over the members in ascending order, odd one bounced to the next
 *     round (analyze_closures__pairwise_grouping, determinise.c:898-1054), each merge done by
 *     combine_result_pair_and_commit (determinise.c:2331-2505).
 * Both steps are pure functions of the epsilon-free NFA; their memo tables only save time.
 * Working out what they compute gives closed forms that need no label-set arithmetic:
 *
 *   leaf     The entries of a single state's result partition the symbols by the SET of
 *            destinations they lead to, listed in descending order of that set read as a bit
 *            vector whose most significant bit is the smallest destination (the greedy sweep
 *            of determinise.c:1180-1290 keeps intersecting with later overlapping groups, i.e.
 *            prefers "in" at the earliest group).
 *   combine  Entries of merge(a, b), with l = the side with fewer entries (a on ties) and r the
 *            other: for each l entry in order, its intersections with r's entries in r order,
 *            then its remainder; then r's remainders in r order.  In other words symbols are
 *            ordered by the pair (index in l, index in r), "no entry" last.
 *
 * So a result is a RANK VECTOR over byte classes (entry index of the class, RN_INF = no
 * transition), a leaf is a dense ranking of destination lists, and a merge is a dense ranking
 * of rank pairs.  One thread evaluates one DFA state's tournament with a binary-counter stack
 * (the bounce rule makes round r's element j the complete subtree over members
 * [j*2^r, (j+1)*2^r) and folds the ragged tail right to left).  The LIFO walk that turns the
 * per-state successor orders into state numbers is inherently sequential (O(edges)) and runs on
 * the host between two kernels.
 *
 * Everything here is plain C++ usable from host and device code: tests/refnum_host.cpp compiles
 * the SAME functions with g++ and checks them against the reference's recorded numbering.
 */
#ifndef FSM_B200_REFNUM_H
#define FSM_B200_REFNUM_H

#include <stdint.h>

#ifdef __CUDACC__
#define RN_HD __host__ __device__ inline
#else
#define RN_HD static inline
#endif

#define RN_INF 0xFFFFu
#define RN_MAX_CLASSES 256u
#define RN_MAX_DEPTH 34u        /* bits(2^32 - 1) + 2 */

/* Order of two destination lists (sorted ascending, duplicates allowed) as signatures:
 * <0 if a is listed first.  At the first differing destination the list holding the smaller one
 * comes first; a proper prefix comes after the longer list. */
RN_HD int
rn_sig_cmp(const uint32_t *a, uint32_t na, const uint32_t *b, uint32_t nb)
{
	uint32_t i = 0, j = 0;
	while (i < na && j < nb) {
		const uint32_t x = a[i], y = b[j];
		if (x != y) return x < y ? -1 : 1;
		do i++; while (i < na && a[i] == x);
		do j++; while (j < nb && b[j] == y);
	}
	if (i >= na && j >= nb) return 0;
	return i >= na ? 1 : -1;
}

/* Leaf: rank vector of one NFA state.  adj_off points at the state's K+1 list offsets into
 * adj_to.  Returns the entry count. */
RN_HD uint32_t
rn_leaf_ranks(const uint64_t *adj_off, const uint32_t *adj_to, uint32_t K, uint16_t *rank)
{
	uint16_t ord[RN_MAX_CLASSES];
	uint32_t m = 0;
	for (uint32_t c = 0; c < K; c++) {
		const uint32_t len = (uint32_t) (adj_off[c + 1] - adj_off[c]);
		if (len == 0) { rank[c] = RN_INF; continue; }
		/* insertion sort; the lists of one state are few and differ early */
		const uint32_t *lc = adj_to + adj_off[c];
		uint32_t p = m;
		while (p > 0) {
			const uint32_t o = ord[p - 1];
			if (rn_sig_cmp(adj_to + adj_off[o], (uint32_t) (adj_off[o + 1] - adj_off[o]), lc, len) <= 0) break;
			ord[p] = ord[p - 1];
			p--;
		}
		ord[p] = (uint16_t) c;
		m++;
	}
	uint32_t r = 0;
	for (uint32_t i = 0; i < m; i++) {
		const uint32_t c = ord[i];
		if (i > 0) {
			const uint32_t o = ord[i - 1];
			if (rn_sig_cmp(adj_to + adj_off[o], (uint32_t) (adj_off[o + 1] - adj_off[o]),
			        adj_to + adj_off[c], (uint32_t) (adj_off[c + 1] - adj_off[c])) != 0) r++;
		}
		rank[c] = (uint16_t) r;
	}
	return m == 0 ? 0 : r + 1;
}

/* Merge: dense ranking of (l[c], r[c]) pairs, RN_INF last, both RN_INF -> RN_INF.
 * out may alias a or b.  Returns the entry count. */
RN_HD uint32_t
rn_combine(const uint16_t *a, uint32_t ma, const uint16_t *b, uint32_t mb, uint32_t K, uint16_t *out)
{
	const uint16_t *l = a, *r = b;
	uint32_t ml = ma, mr = mb;
	if (ma > mb) { l = b; r = a; ml = mb; mr = ma; }
	uint16_t cnt[RN_MAX_CLASSES + 2];
	uint16_t o1[RN_MAX_CLASSES], o2[RN_MAX_CLASSES];

	/* stable counting sort by the r index, then by the l index */
	for (uint32_t i = 0; i <= mr + 1; i++) cnt[i] = 0;
	for (uint32_t c = 0; c < K; c++) { const uint32_t v = r[c] == RN_INF ? mr : r[c]; cnt[v + 1]++; }
	for (uint32_t i = 1; i <= mr + 1; i++) cnt[i] = (uint16_t) (cnt[i] + cnt[i - 1]);
	for (uint32_t c = 0; c < K; c++) { const uint32_t v = r[c] == RN_INF ? mr : r[c]; o1[cnt[v]++] = (uint16_t) c; }
	for (uint32_t i = 0; i <= ml + 1; i++) cnt[i] = 0;
	for (uint32_t c = 0; c < K; c++) { const uint32_t v = l[c] == RN_INF ? ml : l[c]; cnt[v + 1]++; }
	for (uint32_t i = 1; i <= ml + 1; i++) cnt[i] = (uint16_t) (cnt[i] + cnt[i - 1]);
	for (uint32_t i = 0; i < K; i++) { const uint32_t c = o1[i]; const uint32_t v = l[c] == RN_INF ? ml : l[c]; o2[cnt[v]++] = (uint16_t) c; }

	uint32_t rank = 0, prev = 0;
	bool any = false;
	for (uint32_t i = 0; i < K; i++) {
		const uint32_t c = o2[i];
		const uint32_t lc = l[c], rc = r[c];
		if (lc == RN_INF && rc == RN_INF) { out[c] = RN_INF; continue; }
		const uint32_t key = (lc << 16) | rc;
		if (!any) { any = true; prev = key; }
		else if (key != prev) { rank++; prev = key; }
		out[c] = (uint16_t) rank;
	}
	return any ? rank + 1 : 0;
}

/* Stack depth rn_state_rank needs for sets of up to kmax members. */
RN_HD uint32_t
rn_depth_for(uint32_t kmax)
{
	uint32_t bits = 0;
	while (kmax) { bits++; kmax >>= 1; }
	return bits + 1;
}

/* Tournament over one DFA state's members (ascending).  leaf: [nfa states][K] rank vectors,
 * leaf_m: their entry counts; bufs: depth * K scratch.  Returns the entry count; *result points
 * at the final rank vector (inside leaf or bufs). */
RN_HD uint32_t
rn_state_rank(const uint32_t *members, uint32_t k, const uint16_t *leaf, const uint16_t *leaf_m,
	uint32_t K, uint16_t *bufs, const uint16_t **result)
{
	const uint16_t *sp_p[RN_MAX_DEPTH];
	uint16_t sp_m[RN_MAX_DEPTH];
	uint32_t sp = 0;
	for (uint32_t i = 0; i < k; i++) {
		const uint32_t s = members[i];
		sp_p[sp] = leaf + (size_t) s * K;
		sp_m[sp] = leaf_m[s];
		sp++;
		for (uint32_t t = i + 1; (t & 1u) == 0; t >>= 1) {
			uint16_t *dst = bufs + (size_t) (sp - 2) * K;
			sp_m[sp - 2] = (uint16_t) rn_combine(sp_p[sp - 2], sp_m[sp - 2], sp_p[sp - 1], sp_m[sp - 1], K, dst);
			sp_p[sp - 2] = dst;
			sp--;
		}
	}
	while (sp > 1) {
		uint16_t *dst = bufs + (size_t) (sp - 2) * K;
		sp_m[sp - 2] = (uint16_t) rn_combine(sp_p[sp - 2], sp_m[sp - 2], sp_p[sp - 1], sp_m[sp - 1], K, dst);
		sp_p[sp - 2] = dst;
		sp--;
	}
	*result = sp_p[0];
	return sp_m[0];
}

/* Successor order of one DFA state: order[r] = destination of entry r.  trans_row: the state's
 * [K] transition row.  Returns the entry count. */
RN_HD uint32_t
rn_state_order(const uint32_t *members, uint32_t k, const uint16_t *leaf, const uint16_t *leaf_m,
	uint32_t K, uint16_t *bufs, const uint32_t *trans_row, uint32_t *order)
{
	const uint16_t *rk;
	const uint32_t m = rn_state_rank(members, k, leaf, leaf_m, K, bufs, &rk);
	for (uint32_t c = 0; c < K; c++) if (rk[c] != RN_INF) order[rk[c]] = trans_row[c];
	return m;
}

#ifdef __cplusplus
#include <vector>
/* The reference's worklist (determinise.c:118-185): state 0 first; a popped state numbers its
 * not-yet-seen successors in entry order and pushes them; the last pushed is popped next.
 * order: [D][K], order_m: [D].  perm[old] = reference number. */
static inline void
rn_lifo_numbering(uint32_t D, uint32_t K, const uint32_t *order, const uint16_t *order_m, std::vector<uint32_t> &perm)
{
	perm.assign(D, UINT32_MAX);
	if (D == 0) return;
	std::vector<uint32_t> stack;
	uint32_t next = 0, cur = 0;
	perm[0] = next++;
	for (;;) {
		const uint32_t *row = order + (size_t) cur * K;
		for (uint32_t r = 0; r < order_m[cur]; r++) {
			const uint32_t d = row[r];
			if (perm[d] == UINT32_MAX) { perm[d] = next++; stack.push_back(d); }
		}
		if (stack.empty()) break;
		cur = stack.back();
		stack.pop_back();
	}
}
#endif

#endif /* FSM_B200_REFNUM_H */
