/*
 * k2_determinise.cu -- K2: NFA -> DFA subset construction on the GPU.
 *
 * Replaces the reference's fsm_determinise_with_config (src/libfsm/determinise.c:23-335),
 * including the epsilon removal it runs first (fsm_remove_epsilons, src/libfsm/epsilons.c:
 * 121-288; epsilon_closure, src/libfsm/closure.c:130-190).  The reference explores one DFA
 * state at a time from a LIFO worklist and analyses its label groups pairwise with memo
 * tables (determinise.c:898-1054, :1056-1335, :2331-2505).  Here the whole BFS frontier is
 * expanded at once:
 *
 *   K2a  closure     epsilon closure as a dense bit matrix (rows: states with outgoing
 *                    epsilons, columns: epsilon targets), iterated row |= row[target] to the
 *                    fixpoint -- a data-parallel bitset OR-reduction.
 *        adjacency   "edge-set union" of epsilons.c:235-245: for every state s and byte class
 *                    k the destinations of all groups of all closure members, as a CSR keyed
 *                    (s, k).  Byte classes (symbols no label set distinguishes) are derived
 *                    on the host from the 256-bit label sets; only one representative symbol
 *                    per class is expanded (config 5: 27 classes instead of 256 symbols).
 *   K2   expand      one thread per (frontier DFA state, class): gather the members'
 *                    destination lists, sort + unique -> the successor NFA-state set.
 *        intern      open-addressed hash table over set contents (insert-or-find with
 *                    atomicCAS; duplicates inside a round resolved deterministically by the
 *                    smallest candidate index) -> DFA state ids; new sets form the next
 *                    frontier.  interned_state_set_intern_set / map_find / map_add of the
 *                    reference (internedstateset.c:263, determinise.c:486-608).
 *
 * The DFA is the reference's up to state numbering (BFS discovery order here; LIFO worklist
 * plus analysis order there): same sets of NFA states, hence same state count, language,
 * end bits and end-id sets (carried as determinise.c:236-266 / endids.c:782-826 do).
 * With FSM_B200_DET_REFERENCE_NUMBERING the reference's numbering is reproduced as well:
 *        refnum      per-state successor orders from rank vectors (refnum.h), the LIFO walk on
 *                    the host, renumbering on the device before the emit.
 * Integer / set workload: no tensor-core shape anywhere.
 */
#include "k23_common.cuh"
#include "refnum.h"
#include "eager_host.h"

namespace {

constexpr uint32_t CAND_FLAG = 0x80000000u;

thread_local fsm_b200_det_stats tl_stats;


/* ------------------------------------------------------------------ K2a kernels -------- */

struct NfaDev {
	uint32_t n;
	const uint32_t *goff;      /* [n+1] */
	const uint64_t *gsym;      /* [4*G] */
	const uint32_t *gto;       /* [G] */
	const uint32_t *eoff;      /* [n+1] */
	const uint32_t *eto;       /* [E] */
	const uint8_t *is_end;     /* [n] */
};

/* closure bit matrix: seed with the direct epsilon edges */
__global__ void
k2_closure_seed_kernel(NfaDev nfa, const uint32_t *row_of, const uint32_t *col_of, uint32_t W, uint32_t *bits)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= nfa.n) return;
	const uint32_t r = row_of[s];
	if (r == NONE32) return;
	for (uint32_t e = nfa.eoff[s]; e < nfa.eoff[s + 1]; e++) {
		const uint32_t c = col_of[nfa.eto[e]];
		atomicOr(&bits[(size_t) r * W + (c >> 5)], 1u << (c & 31));
	}
}

/* one propagation round: row(s) |= row(t) for every direct epsilon edge s -> t.
 * One block per row; monotone, so concurrent readers of a row being updated are harmless. */
__global__ void
k2_closure_round_kernel(NfaDev nfa, const uint32_t *row_state, const uint32_t *row_of, uint32_t W,
	uint32_t *bits, uint32_t *changed)
{
	const uint32_t r = blockIdx.x;
	const uint32_t s = row_state[r];
	uint32_t *mine = bits + (size_t) r * W;
	bool any = false;
	for (uint32_t e = nfa.eoff[s]; e < nfa.eoff[s + 1]; e++) {
		const uint32_t tr = row_of[nfa.eto[e]];
		if (tr == NONE32 || tr == r) continue;
		const uint32_t *other = bits + (size_t) tr * W;
		for (uint32_t w = threadIdx.x; w < W; w += blockDim.x) {
			const uint32_t o = other[w], m = mine[w];
			if ((o | m) != m) { mine[w] = o | m; any = true; }
		}
	}
	if (any) *changed = 1;
}

/* closure sizes: |{s} U bits(row(s))| */
__global__ void
k2_closure_count_kernel(uint32_t n, const uint32_t *row_of, const uint32_t *col_of, uint32_t W,
	const uint32_t *bits, uint32_t *cnt)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n) return;
	const uint32_t r = row_of[s];
	uint32_t c = 1;
	if (r != NONE32) {
		const uint32_t *row = bits + (size_t) r * W;
		for (uint32_t w = 0; w < W; w++) c += __popc(row[w]);
		const uint32_t sc = col_of[s];
		if (sc != NONE32 && ((row[sc >> 5] >> (sc & 31)) & 1u)) c--;      /* s itself is in its row */
	}
	cnt[s] = c;
}

/* closure members, ascending (columns are numbered in ascending state order) */
__global__ void
k2_closure_fill_kernel(uint32_t n, const uint32_t *row_of, const uint32_t *col_state, uint32_t W,
	const uint32_t *bits, const uint64_t *cl_off, uint32_t *cl_to)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n) return;
	uint64_t o = cl_off[s];
	const uint32_t r = row_of[s];
	bool self_done = false;
	if (r != NONE32) {
		const uint32_t *row = bits + (size_t) r * W;
		for (uint32_t w = 0; w < W; w++) {
			uint32_t m = row[w];
			while (m) {
				const uint32_t b = __ffs(m) - 1;
				m &= m - 1;
				const uint32_t t = col_state[w * 32 + b];
				if (!self_done && t >= s) {
					self_done = true;
					cl_to[o++] = s;
					if (t == s) continue;
				}
				cl_to[o++] = t;
			}
		}
	}
	if (!self_done) cl_to[o++] = s;
}

/* adjacency, pass 1 (count) and pass 2 (fill): thread per state s walks closure(s) x groups */
template <bool FILL>
__global__ void
k2_adjacency_kernel(NfaDev nfa, const uint64_t *cl_off, const uint32_t *cl_to, const uint64_t *gcls,
	uint32_t K, uint32_t *cnt, const uint64_t *adj_off, uint32_t *cursor, uint32_t *adj_to, uint8_t *aend)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= nfa.n) return;
	uint8_t end = 0;
	const uint64_t c0 = cl_off ? cl_off[s] : 0, c1 = cl_off ? cl_off[s + 1] : 1;
	for (uint64_t ci = c0; ci < c1; ci++) {
		const uint32_t e = cl_off ? cl_to[ci] : s;
		end |= nfa.is_end[e];
		for (uint32_t g = nfa.goff[e]; g < nfa.goff[e + 1]; g++) {
			const uint32_t to = nfa.gto[g];
			for (int w = 0; w < 4; w++) {
				uint64_t m = gcls[4 * (size_t) g + w];
				while (m) {
					const uint32_t k = 64 * w + (uint32_t) __ffsll((long long) m) - 1;
					m &= m - 1;
					const size_t key = (size_t) s * K + k;
					if (FILL) adj_to[adj_off[key] + cursor[key]++] = to;
					else cnt[key]++;
				}
			}
		}
	}
	if (!FILL) aend[s] = end;
}

/* sort every (state, class) destination list once, so that the per-candidate gather below
 * concatenates SORTED runs (one long hub run + a few short ones in practice) */
__global__ void
k2_adjacency_sort_kernel(uint64_t nkeys, const uint64_t *adj_off, uint32_t *adj_to)
{
	const uint64_t key = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (key >= nkeys) return;
	uint32_t *a = adj_to + adj_off[key];
	const uint32_t n = (uint32_t) (adj_off[key + 1] - adj_off[key]);
	if (n < 2) return;
	if (n <= 64) {
		for (uint32_t i = 1; i < n; i++) {
			const uint32_t v = a[i];
			uint32_t j = i;
			while (j > 0 && a[j - 1] > v) { a[j] = a[j - 1]; j--; }
			a[j] = v;
		}
		return;
	}
	for (uint32_t start = n / 2; start-- > 0; ) {
		uint32_t root = start;
		for (;;) {
			uint32_t child = 2 * root + 1;
			if (child >= n) break;
			if (child + 1 < n && a[child] < a[child + 1]) child++;
			if (a[root] >= a[child]) break;
			const uint32_t t = a[root]; a[root] = a[child]; a[child] = t;
			root = child;
		}
	}
	for (uint32_t end = n; end-- > 1; ) {
		const uint32_t t0 = a[0]; a[0] = a[end]; a[end] = t0;
		uint32_t root = 0;
		for (;;) {
			uint32_t child = 2 * root + 1;
			if (child >= end) break;
			if (child + 1 < end && a[child] < a[child + 1]) child++;
			if (a[root] >= a[child]) break;
			const uint32_t t = a[root]; a[root] = a[child]; a[child] = t;
			root = child;
		}
	}
}

/* ------------------------------------------------------------------ K2 kernels --------- */

struct Pool {
	const uint64_t *off;       /* [nsets+1] */
	const uint32_t *data;
};

/* E1: upper bound of the successor-set size of candidate (f, k) */
__global__ void
k2_expand_count_kernel(Pool pool, uint32_t fbeg, uint32_t nf, uint32_t K, const uint64_t *adj_off, uint32_t *ub)
{
	const uint64_t c = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= (uint64_t) nf * K) return;
	const uint32_t f = fbeg + (uint32_t) (c / K), k = (uint32_t) (c % K);
	uint64_t sum = 0;
	for (uint64_t i = pool.off[f]; i < pool.off[f + 1]; i++) {
		const size_t key = (size_t) pool.data[i] * K + k;
		sum += adj_off[key + 1] - adj_off[key];
	}
	ub[c] = (uint32_t) sum;
}

/* E2: gather, sort, unique, hash.  One thread per candidate; lists are short for practical
 * NFAs (a handful of members, hub states contribute tens of destinations). */
__global__ void
k2_expand_fill_kernel(Pool pool, uint32_t fbeg, uint32_t nf, uint32_t K, const uint64_t *adj_off,
	const uint32_t *adj_to, const uint64_t *coff, uint32_t *scratch, uint32_t *len, uint64_t *hash)
{
	const uint64_t c = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= (uint64_t) nf * K) return;
	const uint32_t f = fbeg + (uint32_t) (c / K), k = (uint32_t) (c % K);
	uint32_t *a = scratch + coff[c];
	const uint32_t ub = (uint32_t) (coff[c + 1] - coff[c]);
	if (ub == 0) { len[c] = 0; hash[c] = 0; return; }
	uint32_t n = 0;
	for (uint64_t i = pool.off[f]; i < pool.off[f + 1]; i++) {
		const size_t key = (size_t) pool.data[i] * K + k;
		for (uint64_t j = adj_off[key]; j < adj_off[key + 1]; j++) a[n++] = adj_to[j];
	}
	if (n <= 512) {                     /* insertion sort: the input is a few sorted runs */
		for (uint32_t i = 1; i < n; i++) {
			const uint32_t v = a[i];
			uint32_t j = i;
			while (j > 0 && a[j - 1] > v) { a[j] = a[j - 1]; j--; }
			a[j] = v;
		}
	} else {                            /* heap sort */
		for (uint32_t start = n / 2; start-- > 0; ) {
			uint32_t root = start;
			for (;;) {
				uint32_t child = 2 * root + 1;
				if (child >= n) break;
				if (child + 1 < n && a[child] < a[child + 1]) child++;
				if (a[root] >= a[child]) break;
				const uint32_t t = a[root]; a[root] = a[child]; a[child] = t;
				root = child;
			}
		}
		for (uint32_t end = n; end-- > 1; ) {
			const uint32_t t0 = a[0]; a[0] = a[end]; a[end] = t0;
			uint32_t root = 0;
			for (;;) {
				uint32_t child = 2 * root + 1;
				if (child >= end) break;
				if (child + 1 < end && a[child] < a[child + 1]) child++;
				if (a[root] >= a[child]) break;
				const uint32_t t = a[root]; a[root] = a[child]; a[child] = t;
				root = child;
			}
		}
	}
	uint32_t w = 0;
	uint64_t h = 0x243f6a8885a308d3ull;
	for (uint32_t i = 0; i < n; i++) {
		if (w == 0 || a[w - 1] != a[i]) { a[w++] = a[i]; h = mix64(h, a[i]); }
	}
	len[c] = w;
	hash[c] = mix64(h, w);
}

struct Table {
	uint64_t *slots;           /* (tag32 << 32) | value; value: set id, or CAND_FLAG | candidate */
	uint64_t mask;
};

__device__ __forceinline__ bool
set_equal(const uint32_t *a, const uint32_t *b, uint32_t n)
{
	for (uint32_t i = 0; i < n; i++) if (a[i] != b[i]) return false;
	return true;
}

/* I1: insert-or-find.  rep[c] = existing set id, or CAND_FLAG | claimant candidate. */
__global__ void
k2_intern_probe_kernel(Table tab, Pool pool, uint64_t ncand, const uint64_t *coff, const uint32_t *scratch,
	const uint32_t *len, const uint64_t *hash, uint32_t *rep, uint32_t *cand_min, uint64_t *claim_slot)
{
	const uint64_t c = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= ncand) return;
	const uint32_t n = len[c];
	if (n == 0) { rep[c] = NONE32; return; }
	const uint64_t h = hash[c];
	const uint64_t tag = (h >> 32) << 32;
	const uint32_t *mine = scratch + coff[c];
	uint64_t slot = h & tab.mask;
	for (;;) {
		uint64_t v = tab.slots[slot];
		if (v == EMPTY64) {
			const uint64_t want = tag | (uint64_t) (CAND_FLAG | (uint32_t) c);
			const uint64_t prev = atomicCAS((unsigned long long *) &tab.slots[slot], (unsigned long long) EMPTY64,
			    (unsigned long long) want);
			if (prev == EMPTY64) {
				rep[c] = CAND_FLAG | (uint32_t) c;
				claim_slot[c] = slot;
				atomicMin(&cand_min[c], (uint32_t) c);
				return;
			}
			v = prev;
		}
		if ((v & 0xFFFFFFFF00000000ull) == tag) {
			const uint32_t val = (uint32_t) v;
			if (val & CAND_FLAG) {
				const uint32_t j = val & ~CAND_FLAG;
				if (len[j] == n && set_equal(scratch + coff[j], mine, n)) {
					rep[c] = val;
					atomicMin(&cand_min[j], (uint32_t) c);
					return;
				}
			} else {
				const uint64_t o = pool.off[val];
				if (pool.off[val + 1] - o == n && set_equal(pool.data + o, mine, n)) {
					rep[c] = val;
					return;
				}
			}
		}
		slot = (slot + 1) & tab.mask;
	}
}

/* I2: flag the canonical (smallest-index) candidate of every newly claimed set */
__global__ void
k2_intern_flag_kernel(uint64_t ncand, const uint32_t *rep, const uint32_t *cand_min, const uint32_t *len,
	uint32_t *flag, uint32_t *newlen_by_cand)
{
	const uint64_t c = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= ncand) return;
	if (rep[c] == (CAND_FLAG | (uint32_t) c)) {        /* c is a claimant */
		const uint32_t canon = cand_min[c];
		flag[canon] = 1;
		newlen_by_cand[canon] = len[c];
	}
}

/* I3: claimants publish their DFA id in the table and copy the set into the pool */
__global__ void
k2_intern_commit_kernel(Table tab, uint64_t ncand, const uint32_t *rep, const uint32_t *cand_min,
	const uint64_t *rank, const uint64_t *claim_slot, const uint64_t *hash, uint32_t base_id,
	const uint64_t *newoff_by_cand, uint64_t pool_base, const uint64_t *coff, const uint32_t *scratch,
	const uint32_t *len, uint64_t *pool_off, uint32_t *pool_data, uint32_t *newid)
{
	const uint64_t c = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= ncand) return;
	if (rep[c] != (CAND_FLAG | (uint32_t) c)) return;
	const uint32_t canon = cand_min[c];
	const uint32_t id = base_id + (uint32_t) rank[canon];
	newid[c] = id;
	tab.slots[claim_slot[c]] = ((hash[c] >> 32) << 32) | (uint64_t) id;
	const uint64_t dst = pool_base + newoff_by_cand[canon];
	const uint32_t n = len[c];
	const uint32_t *src = scratch + coff[c];
	for (uint32_t i = 0; i < n; i++) pool_data[dst + i] = src[i];
	pool_off[id + 1] = dst + n;
}

/* I4: every candidate resolves its destination id */
__global__ void
k2_intern_resolve_kernel(uint64_t ncand, const uint32_t *rep, const uint32_t *newid, uint32_t *trans_out)
{
	const uint64_t c = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= ncand) return;
	const uint32_t r = rep[c];
	if (r == NONE32) trans_out[c] = NONE32;
	else if (r & CAND_FLAG) trans_out[c] = newid[r & ~CAND_FLAG];
	else trans_out[c] = r;
}

/* rebuild the hash table after growth: thread per existing set */
__global__ void
k2_rehash_kernel(Table tab, Pool pool, uint32_t nsets)
{
	const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
	if (id >= nsets) return;
	uint64_t h = 0x243f6a8885a308d3ull;
	const uint64_t o = pool.off[id], e = pool.off[id + 1];
	for (uint64_t i = o; i < e; i++) h = mix64(h, pool.data[i]);
	h = mix64(h, e - o);
	const uint64_t want = ((h >> 32) << 32) | (uint64_t) id;
	uint64_t slot = h & tab.mask;
	for (;;) {
		const uint64_t prev = atomicCAS((unsigned long long *) &tab.slots[slot], (unsigned long long) EMPTY64,
		    (unsigned long long) want);
		if (prev == EMPTY64) return;
		slot = (slot + 1) & tab.mask;
	}
}

/* ------------------------------------------------------------------ reference numbering */

/* leaf rank vectors (refnum.h): thread per NFA state over its K sorted destination lists */
__global__ void
k2_refnum_leaf_kernel(uint32_t n, uint32_t K, const uint64_t *adj_off, const uint32_t *adj_to, uint16_t *leaf, uint16_t *leaf_m)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n) return;
	leaf_m[s] = (uint16_t) rn_leaf_ranks(adj_off + (size_t) s * K, adj_to, K, leaf + (size_t) s * K);
}

__global__ void
k2_refnum_kmax_kernel(const uint64_t *pool_off, uint32_t D, uint32_t *kmax)
{
	const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t k = d < D ? (uint32_t) (pool_off[d + 1] - pool_off[d]) : 0;
	for (int o = 16; o > 0; o >>= 1) k = max(k, __shfl_xor_sync(0xFFFFFFFFu, k, o));
	if ((threadIdx.x & 31) == 0 && k > 0) atomicMax(kmax, k);
}

/* successor order of every DFA state: thread per state runs the member tournament on a
 * binary-counter stack held in its slice of `scratch` (depth * K rank slots per thread) */
__global__ void
k2_refnum_order_kernel(Pool pool, uint32_t D, uint32_t K, const uint16_t *leaf, const uint16_t *leaf_m,
	const uint32_t *trans, uint16_t *scratch, uint32_t depth, uint32_t *order, uint16_t *order_m)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	uint16_t *bufs = scratch + (size_t) t * depth * K;
	for (uint32_t d = t; d < D; d += gridDim.x * blockDim.x) {
		const uint64_t b = pool.off[d];
		order_m[d] = (uint16_t) rn_state_order(pool.data + b, (uint32_t) (pool.off[d + 1] - b), leaf, leaf_m, K, bufs,
		    trans + (size_t) d * K, order + (size_t) d * K);
	}
}

/* trans2[perm[s]][k] = perm[trans[s][k]] */
__global__ void
k2_refnum_permute_kernel(const uint32_t *trans, uint32_t D, uint32_t K, const uint32_t *perm, uint32_t *trans2)
{
	const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (uint64_t) D * K) return;
	const uint32_t s = (uint32_t) (i / K), k = (uint32_t) (i % K);
	const uint32_t to = trans[i];
	trans2[(size_t) perm[s] * K + k] = to == NONE32 ? NONE32 : perm[to];
}

/* ------------------------------------------------------------------ host: byte classes - */

} // namespace

extern "C" void
fsm_b200_desc_free(struct fsm_b200_owned_desc *d)
{
	if (d == nullptr) return;
	delete static_cast<Owner *>(d->owner);
	memset(d, 0, sizeof *d);
}

extern "C" int
fsm_b200_determinise_stats(struct fsm_b200_det_stats *st)
{
	if (st == nullptr) { errno = EINVAL; return -1; }
	*st = tl_stats;
	return 0;
}

static int
determinise_impl(const struct fsm_b200_desc *nfa, int device, size_t state_limit, unsigned flags,
	struct fsm_b200_owned_desc *out)
{
	if (nfa == nullptr || out == nullptr || (nfa->reserved & ~FSM_B200_DESC_EAGER) != 0 || (flags & ~(unsigned) FSM_B200_DET_REFERENCE_NUMBERING)) {
		set_error("determinise: bad argument");
		errno = EINVAL;
		return -1;
	}
	memset(out, 0, sizeof *out);
	memset(&tl_stats, 0, sizeof tl_stats);
	const auto t_begin = std::chrono::steady_clock::now();
	const uint64_t launches0 = fsm_b200_launch_count(0);
	const uint32_t n = nfa->nstates;
	Owner *own = new (std::nothrow) Owner();
	if (own == nullptr) { errno = ENOMEM; return -1; }
	struct Guard { Owner *o; ~Guard() { delete o; } } guard{ own };

	if (state_limit != 0 && n > state_limit) return 1;                 /* determinise.c:65-68 */
	if (!nfa->hasstart || n == 0) {                                    /* determinise.c:88-91 */
		own->group_off.assign(1, 0); own->endid_off.assign(1, 0);
		out->desc.nstates = 0; out->desc.hasstart = 0;
		out->desc.group_off = own->group_off.data(); out->desc.endid_off = own->endid_off.data();
		out->owner = own; guard.o = nullptr;
		return 0;
	}
	const uint64_t G = nfa->group_off[n];
	const uint64_t E = nfa->eps_off ? nfa->eps_off[n] : 0;
	if (G >= (1ull << 31) || E >= (1ull << 31) || nfa->start >= n) {
		set_error("determinise: automaton too large or bad start");
		errno = EINVAL;
		return -1;
	}
	for (uint64_t g = 0; g < G; g++) if (nfa->group_to[g] >= n) { set_error("determinise: edge out of range"); errno = EINVAL; return -1; }
	for (uint64_t e = 0; e < E; e++) if (nfa->eps_to[e] >= n) { set_error("determinise: epsilon out of range"); errno = EINVAL; return -1; }

	uint8_t class_of[256], rep[256];
	const uint32_t K = G > 0 ? byte_classes(nfa, G, class_of, rep) : 1;
	if (G == 0) { rep[0] = 0; memset(class_of, 0, sizeof class_of); }

	CK(cudaSetDevice(device));
	{
		cudaMemPool_t mp;
		uint64_t keep_all = UINT64_MAX;
		if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) {
			cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep_all);
		}
	}
	cudaStream_t st;
	(void) cudaGetLastError();          /* CK_SYNC reports launch failures of THIS call only */
	CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	/* declared before every DBuf so that it is destroyed after their cudaFreeAsync calls */
	struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamSynchronize(s); cudaStreamDestroy(s); } } sguard{ st };
	Scanner scan; scan.st = st;

	/* ---- upload the NFA ---- */
	std::vector<uint32_t> h_goff(n + 1), h_eoff(n + 1, 0);
	for (uint32_t s = 0; s <= n; s++) h_goff[s] = (uint32_t) nfa->group_off[s];
	if (nfa->eps_off) for (uint32_t s = 0; s <= n; s++) h_eoff[s] = (uint32_t) nfa->eps_off[s];
	DBuf<uint32_t> d_goff, d_gto, d_eoff, d_eto;
	DBuf<uint64_t> d_gsym, d_gcls;
	DBuf<uint8_t> d_isend, d_rep, d_aend;
	if (d_goff.reserve(n + 1, false, st) || d_gto.reserve(G + 1, false, st) || d_eoff.reserve(n + 1, false, st) ||
	    d_eto.reserve(E + 1, false, st) || d_gsym.reserve(4 * G + 4, false, st) || d_gcls.reserve(4 * G + 4, false, st) ||
	    d_isend.reserve(n, false, st) || d_rep.reserve(256, false, st) || d_aend.reserve(n, false, st)) return -1;
	CK(cudaMemcpyAsync(d_goff.p, h_goff.data(), (n + 1) * 4, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_eoff.p, h_eoff.data(), (n + 1) * 4, cudaMemcpyHostToDevice, st));
	if (G) {
		CK(cudaMemcpyAsync(d_gto.p, nfa->group_to, G * 4, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(d_gsym.p, nfa->group_symbols, G * 32, cudaMemcpyHostToDevice, st));
	}
	if (E) CK(cudaMemcpyAsync(d_eto.p, nfa->eps_to, E * 4, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_isend.p, nfa->is_end, n, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_rep.p, rep, 256, cudaMemcpyHostToDevice, st));
	NfaDev dn{ n, d_goff.p, d_gsym.p, d_gto.p, d_eoff.p, d_eto.p, d_isend.p };
	if (G) { k2_group_classmask_kernel<<<blocks_for(G), 256, 0, st>>>(d_gsym.p, (uint32_t) G, d_rep.p, K, d_gcls.p); count_launch(); }

	/* ---- K2a: epsilon closure ---- */
	const auto t_cl = std::chrono::steady_clock::now();
	DBuf<uint64_t> d_cloff;
	DBuf<uint32_t> d_clto;
	std::vector<uint64_t> h_cloff;
	std::vector<uint32_t> h_clto;
	bool have_closure = false;
	if (E > 0) {
		std::vector<uint32_t> row_of(n, NONE32), col_of(n, NONE32), row_state, col_state;
		for (uint32_t s = 0; s < n; s++) if (h_eoff[s + 1] > h_eoff[s]) { row_of[s] = (uint32_t) row_state.size(); row_state.push_back(s); }
		std::vector<uint8_t> is_target(n, 0);
		for (uint64_t e = 0; e < E; e++) is_target[nfa->eps_to[e]] = 1;
		for (uint32_t s = 0; s < n; s++) if (is_target[s]) { col_of[s] = (uint32_t) col_state.size(); col_state.push_back(s); }
		const uint32_t R = (uint32_t) row_state.size(), Cn = (uint32_t) col_state.size();
		const uint32_t W = (Cn + 31) / 32;
		const uint64_t bit_words = (uint64_t) R * W;
		if (bit_words * 4 > (64ull << 30)) { set_error("determinise: epsilon closure matrix too large (%u x %u)", R, Cn); errno = ENOMEM; return -1; }
		DBuf<uint32_t> d_rowof, d_colof, d_rowstate, d_colstate, d_bits, d_changed, d_cnt;
		if (d_rowof.reserve(n, false, st) || d_colof.reserve(n, false, st) || d_rowstate.reserve(R + 1, false, st) ||
		    d_colstate.reserve((size_t) W * 32 + 32, false, st) || d_bits.reserve(bit_words + 1, false, st) ||
		    d_changed.reserve(1, false, st) || d_cnt.reserve(n + 1, false, st)) return -1;
		col_state.resize((size_t) W * 32, NONE32);
		CK(cudaMemcpyAsync(d_rowof.p, row_of.data(), n * 4, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(d_colof.p, col_of.data(), n * 4, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(d_rowstate.p, row_state.data(), R * 4, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(d_colstate.p, col_state.data(), (size_t) W * 32 * 4, cudaMemcpyHostToDevice, st));
		CK(cudaMemsetAsync(d_bits.p, 0, bit_words * 4, st));
		k2_closure_seed_kernel<<<blocks_for(n), 256, 0, st>>>(dn, d_rowof.p, d_colof.p, W, d_bits.p); count_launch();
		for (uint32_t round = 0; round < n + 2; round++) {
			uint32_t changed = 0;
			CK(cudaMemsetAsync(d_changed.p, 0, 4, st));
			k2_closure_round_kernel<<<R, 128, 0, st>>>(dn, d_rowstate.p, d_rowof.p, W, d_bits.p, d_changed.p); count_launch();
			CK(cudaMemcpyAsync(&changed, d_changed.p, 4, cudaMemcpyDeviceToHost, st));
			CK_SYNC(st);
			if (!changed) break;
		}
		k2_closure_count_kernel<<<blocks_for(n), 256, 0, st>>>(n, d_rowof.p, d_colof.p, W, d_bits.p, d_cnt.p); count_launch();
		if (d_cloff.reserve(n + 2, false, st)) return -1;
		if (scan.run<uint32_t>(d_cnt.p, d_cloff.p, n) != 0) return -1;
		h_cloff.resize(n + 1);
		CK(cudaMemcpyAsync(h_cloff.data(), d_cloff.p, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
		CK_SYNC(st);
		const uint64_t total = h_cloff[n];
		if (d_clto.reserve(total + 1, false, st)) return -1;
		k2_closure_fill_kernel<<<blocks_for(n), 256, 0, st>>>(n, d_rowof.p, d_colstate.p, W, d_bits.p, d_cloff.p, d_clto.p); count_launch();
		h_clto.resize(total);
		CK(cudaMemcpyAsync(h_clto.data(), d_clto.p, total * 4, cudaMemcpyDeviceToHost, st));
		CK_SYNC(st);
		have_closure = true;
	}
	tl_stats.ms_closure = ms_since(t_cl);

	/* ---- adjacency (edge-set union over closures), keyed (state, class) ---- */
	const uint64_t NK = (uint64_t) n * K;
	DBuf<uint32_t> d_cnt2, d_cursor, d_adjto;
	DBuf<uint64_t> d_adjoff;
	if (d_cnt2.reserve(NK + 1, false, st) || d_cursor.reserve(NK + 1, false, st) || d_adjoff.reserve(NK + 2, false, st)) return -1;
	CK(cudaMemsetAsync(d_cnt2.p, 0, (NK + 1) * 4, st));
	CK(cudaMemsetAsync(d_cursor.p, 0, (NK + 1) * 4, st));
	k2_adjacency_kernel<false><<<blocks_for(n, 128), 128, 0, st>>>(dn, have_closure ? d_cloff.p : nullptr, d_clto.p, d_gcls.p, K,
	    d_cnt2.p, nullptr, nullptr, nullptr, d_aend.p); count_launch();
	if (scan.run<uint32_t>(d_cnt2.p, d_adjoff.p, NK) != 0) return -1;
	uint64_t adj_total = 0;
	CK(cudaMemcpyAsync(&adj_total, d_adjoff.p + NK, 8, cudaMemcpyDeviceToHost, st));
	CK_SYNC(st);
	if (d_adjto.reserve(adj_total + 1, false, st)) return -1;
	k2_adjacency_kernel<true><<<blocks_for(n, 128), 128, 0, st>>>(dn, have_closure ? d_cloff.p : nullptr, d_clto.p, d_gcls.p, K,
	    nullptr, d_adjoff.p, d_cursor.p, d_adjto.p, nullptr); count_launch();
	k2_adjacency_sort_kernel<<<blocks_for(NK), 256, 0, st>>>(NK, d_adjoff.p, d_adjto.p); count_launch();

	/* ---- K2: frontier-batched subset construction ---- */
	const auto t_exp = std::chrono::steady_clock::now();
	DBuf<uint64_t> d_pooloff, d_slots, d_coff, d_hash, d_claim, d_rank, d_newoff;
	DBuf<uint32_t> d_pooldata, d_trans, d_ub, d_scratch, d_len, d_rep2, d_candmin, d_flag, d_newlen, d_newid;
	uint64_t tab_cap = 1ull << 16;
	uint32_t nsets = 1;
	uint64_t pool_used = 1;
	if (d_pooloff.reserve(1 << 16, false, st) || d_pooldata.reserve(1 << 18, false, st) || d_slots.reserve(tab_cap, false, st) ||
	    d_trans.reserve((size_t) (1 << 12) * K, false, st)) return -1;
	{
		const uint64_t off0[2] = { 0, 1 };
		const uint32_t start = nfa->start;
		CK(cudaMemcpyAsync(d_pooloff.p, off0, 16, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(d_pooldata.p, &start, 4, cudaMemcpyHostToDevice, st));
		k2_fill_u64_kernel<<<blocks_for(tab_cap), 256, 0, st>>>(d_slots.p, EMPTY64, tab_cap); count_launch();
		Table tab{ d_slots.p, tab_cap - 1 };
		k2_rehash_kernel<<<1, 32, 0, st>>>(tab, Pool{ d_pooloff.p, d_pooldata.p }, 1); count_launch();
	}
	uint32_t fbeg = 0;
	uint64_t rounds = 0;
	const uint32_t max_nf = std::max(1u, (1u << 28) / K);   /* candidate indices must fit 31 bits */
	while (fbeg < nsets) {
		rounds++;
		const uint32_t nf = std::min(nsets - fbeg, max_nf);
		const uint32_t fend = fbeg + nf;
		const uint64_t ncand = (uint64_t) nf * K;
		if (d_ub.reserve(ncand + 1, false, st) || d_coff.reserve(ncand + 2, false, st) || d_len.reserve(ncand, false, st) ||
		    d_hash.reserve(ncand, false, st) || d_rep2.reserve(ncand, false, st) || d_candmin.reserve(ncand, false, st) ||
		    d_claim.reserve(ncand, false, st) || d_flag.reserve(ncand + 1, false, st) || d_newlen.reserve(ncand + 1, false, st) ||
		    d_rank.reserve(ncand + 2, false, st) || d_newoff.reserve(ncand + 2, false, st) || d_newid.reserve(ncand, false, st) ||
		    d_trans.reserve((size_t) fend * K, true, st)) return -1;
		Pool pool{ d_pooloff.p, d_pooldata.p };
		k2_expand_count_kernel<<<blocks_for(ncand), 256, 0, st>>>(pool, fbeg, nf, K, d_adjoff.p, d_ub.p); count_launch();
		if (scan.run<uint32_t>(d_ub.p, d_coff.p, ncand) != 0) return -1;
		uint64_t scratch_total = 0;
		CK(cudaMemcpyAsync(&scratch_total, d_coff.p + ncand, 8, cudaMemcpyDeviceToHost, st));
		CK_SYNC(st);
		if (d_scratch.reserve(scratch_total + 1, false, st)) return -1;
		k2_expand_fill_kernel<<<blocks_for(ncand, 128), 128, 0, st>>>(pool, fbeg, nf, K, d_adjoff.p, d_adjto.p, d_coff.p,
		    d_scratch.p, d_len.p, d_hash.p); count_launch();

		/* keep the table at most half full even if every candidate is new */
		if ((uint64_t) nsets + ncand > tab_cap / 2) {
			while ((uint64_t) nsets + ncand > tab_cap / 2) tab_cap <<= 1;
			if (d_slots.reserve(tab_cap, false, st)) return -1;
			k2_fill_u64_kernel<<<blocks_for(tab_cap), 256, 0, st>>>(d_slots.p, EMPTY64, tab_cap); count_launch();
			k2_rehash_kernel<<<blocks_for(nsets), 256, 0, st>>>(Table{ d_slots.p, tab_cap - 1 }, pool, nsets); count_launch();
		}
		Table tab{ d_slots.p, tab_cap - 1 };
		k2_fill_u32_kernel<<<blocks_for(ncand), 256, 0, st>>>(d_candmin.p, NONE32, ncand); count_launch();
		CK(cudaMemsetAsync(d_flag.p, 0, (ncand + 1) * 4, st));
		CK(cudaMemsetAsync(d_newlen.p, 0, (ncand + 1) * 4, st));
		k2_intern_probe_kernel<<<blocks_for(ncand, 128), 128, 0, st>>>(tab, pool, ncand, d_coff.p, d_scratch.p, d_len.p, d_hash.p,
		    d_rep2.p, d_candmin.p, d_claim.p); count_launch();
		k2_intern_flag_kernel<<<blocks_for(ncand), 256, 0, st>>>(ncand, d_rep2.p, d_candmin.p, d_len.p, d_flag.p, d_newlen.p); count_launch();
		if (scan.run<uint32_t>(d_flag.p, d_rank.p, ncand) != 0) return -1;
		if (scan.run<uint32_t>(d_newlen.p, d_newoff.p, ncand) != 0) return -1;
		uint64_t nnew = 0, newdata = 0;
		CK(cudaMemcpyAsync(&nnew, d_rank.p + ncand, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(&newdata, d_newoff.p + ncand, 8, cudaMemcpyDeviceToHost, st));
		CK_SYNC(st);
		/* determinise.c:166-169: adding a state fails when the count BEFORE adding exceeds
		 * the limit, i.e. at most limit+1 states can ever exist */
		if (state_limit != 0 && (uint64_t) nsets + nnew > (uint64_t) state_limit + 1) return 1;
		if ((uint64_t) nsets + nnew >= (1ull << 31)) { set_error("determinise: too many DFA states"); errno = ENOMEM; return -1; }
		if (d_pooloff.reserve(nsets + nnew + 2, true, st) || d_pooldata.reserve(pool_used + newdata + 1, true, st)) return -1;
		k2_intern_commit_kernel<<<blocks_for(ncand, 128), 128, 0, st>>>(tab, ncand, d_rep2.p, d_candmin.p, d_rank.p, d_claim.p,
		    d_hash.p, nsets, d_newoff.p, pool_used, d_coff.p, d_scratch.p, d_len.p, d_pooloff.p, d_pooldata.p, d_newid.p); count_launch();
		k2_intern_resolve_kernel<<<blocks_for(ncand), 256, 0, st>>>(ncand, d_rep2.p, d_newid.p, d_trans.p + (size_t) fbeg * K); count_launch();
		fbeg = fend;
		nsets += (uint32_t) nnew;
		pool_used += newdata;
	}
	CK_SYNC(st);
	tl_stats.ms_expand = ms_since(t_exp);
	tl_stats.rounds = rounds;

	const uint32_t D = nsets;

	/* ---- optional: the reference's state numbering (refnum.h) ----
	 * leaf ranks and per-state successor orders on the device, the LIFO worklist walk
	 * (determinise.c:118-185, sequential by nature) on the host, the renumbering on the
	 * device again; the emit below then runs on the renumbered table. */
	const uint32_t *trans_emit = d_trans.p;
	std::vector<uint32_t> perm, inv;
	DBuf<uint32_t> d_trans2;
	if (flags & FSM_B200_DET_REFERENCE_NUMBERING) {
		const auto t_num = std::chrono::steady_clock::now();
		DBuf<uint16_t> d_leaf, d_leafm, d_rnscratch, d_orderm;
		DBuf<uint32_t> d_order, d_kmax, d_perm;
		if (d_leaf.reserve(NK + 1, false, st) || d_leafm.reserve(n + 1, false, st) || d_kmax.reserve(1, false, st) ||
		    d_order.reserve((size_t) D * K + 1, false, st) || d_orderm.reserve(D + 1, false, st) ||
		    d_perm.reserve(D + 1, false, st) || d_trans2.reserve((size_t) D * K + 1, false, st)) return -1;
		k2_refnum_leaf_kernel<<<blocks_for(n, 128), 128, 0, st>>>(n, K, d_adjoff.p, d_adjto.p, d_leaf.p, d_leafm.p); count_launch();
		CK(cudaMemsetAsync(d_kmax.p, 0, 4, st));
		k2_refnum_kmax_kernel<<<blocks_for(D), 256, 0, st>>>(d_pooloff.p, D, d_kmax.p); count_launch();
		uint32_t kmax = 0;
		CK(cudaMemcpyAsync(&kmax, d_kmax.p, 4, cudaMemcpyDeviceToHost, st));
		CK_SYNC(st);
		const uint32_t depth = rn_depth_for(std::max(kmax, 1u));
		const uint32_t nblk = std::min<uint32_t>(blocks_for(D, 128), 148u * 8u);
		if (d_rnscratch.reserve((size_t) nblk * 128 * depth * K + 1, false, st)) return -1;
		k2_refnum_order_kernel<<<nblk, 128, 0, st>>>(Pool{ d_pooloff.p, d_pooldata.p }, D, K, d_leaf.p, d_leafm.p, d_trans.p,
		    d_rnscratch.p, depth, d_order.p, d_orderm.p); count_launch();
		std::vector<uint32_t> h_order((size_t) D * K);
		std::vector<uint16_t> h_orderm(D);
		CK(cudaMemcpyAsync(h_order.data(), d_order.p, (size_t) D * K * 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(h_orderm.data(), d_orderm.p, (size_t) D * 2, cudaMemcpyDeviceToHost, st));
		CK_SYNC(st);
		for (uint32_t s = 0; s < D; s++) {
			if (h_orderm[s] > K) { set_error("determinise: numbering: bad entry count"); errno = EIO; return -1; }
			for (uint32_t r = 0; r < h_orderm[s]; r++) {
				if (h_order[(size_t) s * K + r] >= D) { set_error("determinise: numbering: bad successor"); errno = EIO; return -1; }
			}
		}
		rn_lifo_numbering(D, K, h_order.data(), h_orderm.data(), perm);
		inv.assign(D, NONE32);
		for (uint32_t s = 0; s < D; s++) {
			if (perm[s] >= D || inv[perm[s]] != NONE32) { set_error("determinise: numbering: not a permutation"); errno = EIO; return -1; }
			inv[perm[s]] = s;
		}
		CK(cudaMemcpyAsync(d_perm.p, perm.data(), (size_t) D * 4, cudaMemcpyHostToDevice, st));
		k2_refnum_permute_kernel<<<blocks_for((uint64_t) D * K), 256, 0, st>>>(d_trans.p, D, K, d_perm.p, d_trans2.p); count_launch();
		CK_SYNC(st);   /* perm (host memory) is read by the copy above */
		trans_emit = d_trans2.p;
		tl_stats.ms_numbering = ms_since(t_num);
	}

	/* ---- emit: groups on the device; end bits + end ids on the host (O(pool)) ---- */
	const auto t_emit = std::chrono::steady_clock::now();
	uint64_t class_mask[256][4];
	memset(class_mask, 0, sizeof class_mask);
	for (int c = 0; c < 256; c++) class_mask[class_of[c]][c >> 6] |= 1ull << (c & 63);
	DBuf<uint64_t> d_cmask, d_ogoff, d_ogsym;
	DBuf<uint32_t> d_ng, d_ogto;
	if (d_cmask.reserve(1024, false, st) || d_ng.reserve(D + 1, false, st) || d_ogoff.reserve(D + 2, false, st)) return -1;
	CK(cudaMemcpyAsync(d_cmask.p, class_mask, sizeof class_mask, cudaMemcpyHostToDevice, st));
	k2_emit_count_kernel<<<blocks_for(D, 128), 128, 0, st>>>(trans_emit, D, K, d_ng.p); count_launch();
	if (scan.run<uint32_t>(d_ng.p, d_ogoff.p, D) != 0) return -1;
	own->group_off.assign(D + 1, 0);
	CK(cudaMemcpyAsync(own->group_off.data(), d_ogoff.p, (D + 1) * 8, cudaMemcpyDeviceToHost, st));
	CK_SYNC(st);
	const uint64_t NG = own->group_off[D];
	if (d_ogto.reserve(NG + 1, false, st) || d_ogsym.reserve(4 * NG + 4, false, st)) return -1;
	k2_emit_fill_kernel<<<blocks_for(D, 128), 128, 0, st>>>(trans_emit, D, K, d_cmask.p, d_ogoff.p, d_ogto.p, d_ogsym.p); count_launch();
	/* the groups (config 5: 94 MB) and the state-set pool land in pinned pool blocks whose pages are already
	 * faulted in: fresh vector storage cost 36-50 ms of page faults per call, more than every kernel together */
	own->groups_alloc(NG);
	PoolTemp t_pooldata, t_pooloff;
	std::vector<uint32_t> v_pooldata;
	std::vector<uint64_t> v_pooloff;
	uint32_t *h_pooldata; uint64_t *h_pooloff;
	if (t_pooldata.get((size_t) pool_used * 4 + 4) && t_pooloff.get((size_t) (D + 1) * 8)) {
		h_pooldata = static_cast<uint32_t *>(t_pooldata.p); h_pooloff = static_cast<uint64_t *>(t_pooloff.p);
	} else {
		v_pooldata.resize(pool_used + 1); v_pooloff.resize(D + 1);
		h_pooldata = v_pooldata.data(); h_pooloff = v_pooloff.data();
	}
	std::vector<uint8_t> h_aend(n);
	if (NG) {
		CK(cudaMemcpyAsync(own->gto(), d_ogto.p, NG * 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(own->gsym(), d_ogsym.p, NG * 32, cudaMemcpyDeviceToHost, st));
	}
	CK(cudaMemcpyAsync(h_pooloff, d_pooloff.p, (D + 1) * 8, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(h_pooldata, d_pooldata.p, pool_used * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(h_aend.data(), d_aend.p, n, cudaMemcpyDeviceToHost, st));
	CK_SYNC(st);

	own->is_end.assign(D, 0);
	own->endid_off.assign(D + 1, 0);
	std::vector<uint32_t> ids, xacc;
	const uint64_t *xoff = nullptr; const uint32_t *xids = nullptr;
	const bool has_eager = eagerhost::eh_get(nfa, &xoff, &xids);
	if (has_eager) own->eager_off.assign(D + 1, 0);
	for (uint32_t s = 0; s < D; s++) {
		/* end bit + end ids: determinise.c:236-266 over the epsilon-folded members */
		ids.clear();
		bool end = false;
		const uint32_t src = inv.empty() ? s : inv[s];    /* pool index of output state s */
		for (uint64_t i = h_pooloff[src]; i < h_pooloff[src + 1]; i++) {
			const uint32_t m = h_pooldata[i];
			if (!h_aend[m]) continue;
			end = true;
			if (nfa->endid_off == nullptr) continue;
			const uint64_t c0 = have_closure ? h_cloff[m] : 0, c1 = have_closure ? h_cloff[m + 1] : 1;
			for (uint64_t ci = c0; ci < c1; ci++) {
				const uint32_t e = have_closure ? h_clto[ci] : m;
				if (!nfa->is_end[e]) continue;
				for (uint64_t q = nfa->endid_off[e]; q < nfa->endid_off[e + 1]; q++) ids.push_back(nfa->endids[q]);
			}
		}
		own->is_end[s] = end ? 1 : 0;
		if (ids.size() > 1) {
			std::sort(ids.begin(), ids.end());
			ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
		}
		own->endids.insert(own->endids.end(), ids.begin(), ids.end());
		own->endid_off[s + 1] = own->endids.size();
		if (has_eager) {
			/* eager outputs: the ids of every state in the closure of every member
			 * (epsilons.c:221-253, determinise.c:2614-2636) */
			xacc.clear();
			for (uint64_t i = h_pooloff[src]; i < h_pooloff[src + 1]; i++) {
				const uint32_t m = h_pooldata[i];
				if (have_closure) eagerhost::eh_union_over(h_clto.begin() + h_cloff[m], h_clto.begin() + h_cloff[m + 1], xoff, xids, xacc);
				else eagerhost::eh_union_over(&m, &m + 1, xoff, xids, xacc);
			}
			eagerhost::eh_sort_unique(xacc);
			own->eager_ids.insert(own->eager_ids.end(), xacc.begin(), xacc.end());
			own->eager_off[s + 1] = own->eager_ids.size();
		}
	}
	if (has_eager && own->eager_ids.empty()) own->eager_off.clear();
	if (NG == 0) { own->group_to.assign(1, 0); own->group_sym.assign(4, 0); }
	if (own->endids.empty()) own->endids.push_back(0);
	tl_stats.ms_emit = ms_since(t_emit);

	out->desc.nstates = D;
	out->desc.start = 0;
	out->desc.hasstart = 1;
	out->desc.is_end = own->is_end.data();
	out->desc.group_off = own->group_off.data();
	out->desc.group_symbols = own->gsym();
	out->desc.group_to = own->gto();
	out->desc.eps_off = nullptr;
	out->desc.eps_to = nullptr;
	out->desc.endid_off = own->endid_off.data();
	out->desc.endids = own->endids.data();
	out->owner = own;
	guard.o = nullptr;
	tl_stats.ms_total = ms_since(t_begin);
	tl_stats.dfa_states = D;
	tl_stats.dfa_groups = own->group_off[D];
	tl_stats.kernel_launches = fsm_b200_launch_count(0) - launches0;
	return 0;
}

extern "C" int
fsm_b200_owned_desc_eager(const struct fsm_b200_owned_desc *d, const uint64_t **eager_off, const uint32_t **eager_ids)
{
	if (d == nullptr || eager_off == nullptr || eager_ids == nullptr) { errno = EINVAL; return -1; }
	const Owner *own = static_cast<const Owner *>(d->owner);
	const bool has = own != nullptr && !own->eager_off.empty();
	*eager_off = has ? own->eager_off.data() : nullptr;
	*eager_ids = has ? own->eager_ids.data() : nullptr;
	return 0;
}

extern "C" int
fsm_b200_determinise_ex(const struct fsm_b200_desc *nfa, int device, size_t state_limit, unsigned flags,
	struct fsm_b200_owned_desc *out)
{
	return determinise_impl(nfa, device, state_limit, flags, out);
}

extern "C" int
fsm_b200_determinise(const struct fsm_b200_desc *nfa, int device, size_t state_limit,
	struct fsm_b200_owned_desc *out)
{
	const char *e = getenv("FSM_B200_DET_NUMBERING");
	const unsigned flags = (e != nullptr && strcmp(e, "reference") == 0) ? FSM_B200_DET_REFERENCE_NUMBERING : 0u;
	return determinise_impl(nfa, device, state_limit, flags, out);
}
