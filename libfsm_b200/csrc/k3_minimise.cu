/*
 * k3_minimise.cu -- K3: DFA minimisation on the GPU.
 *
 * Replaces the partition refinement of the reference's fsm_minimise
 * (src/libfsm/minimise.c:74-190: fsm_trim FSM_TRIM_START_AND_END_REACHABLE, then
 * build_minimised_mapping :252-, split_ecs_by_end_metadata :733-, then fsm_consolidate).
 * The reference refines one equivalence class at a time, label by label, over linked lists.
 * Here every state is refined at once (Moore):
 *
 *   trans      dense [state][byte class] successor array built from the edge groups (byte
 *              classes from the label sets, as K2);
 *   trim       forward reachability from the start and backward reachability from the end
 *              states as frontier sweeps over `trans` until nothing changes;
 *   refine     per round, one thread per state hashes its signature (own class, classes of
 *              its successors per byte class), inserts it into an open-addressed table
 *              (atomicCAS claim, content compare on tag match) and takes the smallest state
 *              of its signature group as canonical; a scan ranks the canonical states ->
 *              new class ids.  Classes only ever split, so the loop ends when the class
 *              count stops growing.
 *   emit       one state per class (numbered by smallest member), groups by destination.
 *
 * End states start in classes by end-id set (minimise.c:733-: states with different end ids
 * are never merged).  The minimal DFA is unique up to numbering, so the result is isomorphic
 * to the reference's on the reference's own pipeline inputs (tests compare canonical forms).
 */
#include "k23_common.cuh"
#include "eager_host.h"

namespace {

/* trans[s][k] from the groups: thread per state */
__global__ void
k3_build_trans_kernel(uint32_t n, const uint32_t *goff, const uint32_t *gto, const uint64_t *gcls, uint32_t K, uint32_t *trans)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n) return;
	for (uint32_t k = 0; k < K; k++) trans[(size_t) s * K + k] = NONE32;
	for (uint32_t g = goff[s]; g < goff[s + 1]; g++) {
		const uint32_t to = gto[g];
		for (int w = 0; w < 4; w++) {
			uint64_t m = gcls[4 * (size_t) g + w];
			while (m) {
				const uint32_t k = 64 * w + (uint32_t) __ffsll((long long) m) - 1;
				m &= m - 1;
				trans[(size_t) s * K + k] = to;
			}
		}
	}
}

/* forward sweep: successors of reached states become reached */
__global__ void
k3_reach_kernel(uint32_t n, uint32_t K, const uint32_t *trans, uint8_t *reach, uint32_t *changed)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n || !reach[s]) return;
	bool any = false;
	for (uint32_t k = 0; k < K; k++) {
		const uint32_t t = trans[(size_t) s * K + k];
		if (t != NONE32 && !reach[t]) { reach[t] = 1; any = true; }
	}
	if (any) *changed = 1;
}

/* backward sweep: a state with a successor that reaches an end reaches an end */
__global__ void
k3_coreach_kernel(uint32_t n, uint32_t K, const uint32_t *trans, uint8_t *co, uint32_t *changed)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n || co[s]) return;
	for (uint32_t k = 0; k < K; k++) {
		const uint32_t t = trans[(size_t) s * K + k];
		if (t != NONE32 && co[t]) { co[s] = 1; *changed = 1; return; }
	}
}

__global__ void
k3_keep_kernel(uint32_t n, const uint8_t *reach, const uint8_t *co, uint32_t *keep)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s < n) keep[s] = (reach[s] && co[s]) ? 1u : 0u;
}

/* compact: kept state s -> index newid[s]; successors outside the kept set become NONE */
__global__ void
k3_compact_kernel(uint32_t n, uint32_t K, const uint32_t *trans, const uint32_t *keep, const uint64_t *newid,
	const uint32_t *cls0, uint32_t *ktrans, uint32_t *kcls, uint32_t *korig)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n || !keep[s]) return;
	const uint32_t i = (uint32_t) newid[s];
	for (uint32_t k = 0; k < K; k++) {
		const uint32_t t = trans[(size_t) s * K + k];
		ktrans[(size_t) i * K + k] = (t != NONE32 && keep[t]) ? (uint32_t) newid[t] : NONE32;
	}
	kcls[i] = cls0[s];
	korig[i] = s;
}

__device__ __forceinline__ uint32_t
sig_at(const uint32_t *cls, const uint32_t *ktrans, uint32_t K, uint32_t i, uint32_t k)
{
	/* k == 0: own class; k >= 1: class of the successor on byte class k-1 */
	if (k == 0) return cls[i];
	const uint32_t t = ktrans[(size_t) i * K + (k - 1)];
	return t == NONE32 ? NONE32 : cls[t];
}

struct RTable { uint64_t *slots; uint64_t mask; };

/* one refinement round, part 1: insert-or-find the signature; rep[i] = claimant state */
__global__ void
k3_refine_probe_kernel(RTable tab, uint32_t m, uint32_t K, const uint32_t *cls, const uint32_t *ktrans,
	uint32_t *rep, uint32_t *cmin)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	uint64_t h = 0x243f6a8885a308d3ull;
	for (uint32_t k = 0; k <= K; k++) h = mix64(h, sig_at(cls, ktrans, K, i, k));
	const uint64_t tag = (h >> 32) << 32;
	uint64_t slot = h & tab.mask;
	for (;;) {
		uint64_t v = tab.slots[slot];
		if (v == EMPTY64) {
			const uint64_t prev = atomicCAS((unsigned long long *) &tab.slots[slot], (unsigned long long) EMPTY64,
			    (unsigned long long) (tag | i));
			if (prev == EMPTY64) { rep[i] = i; atomicMin(&cmin[i], i); return; }
			v = prev;
		}
		if ((v & 0xFFFFFFFF00000000ull) == tag) {
			const uint32_t j = (uint32_t) v;
			bool same = true;
			for (uint32_t k = 0; k <= K && same; k++) same = sig_at(cls, ktrans, K, i, k) == sig_at(cls, ktrans, K, j, k);
			if (same) { rep[i] = j; atomicMin(&cmin[j], i); return; }
		}
		slot = (slot + 1) & tab.mask;
	}
}

/* part 2: the smallest state of every signature group is its canonical member */
__global__ void
k3_refine_flag_kernel(uint32_t m, const uint32_t *rep, const uint32_t *cmin, uint32_t *flag)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < m && rep[i] == i) flag[cmin[i]] = 1;
}

/* part 3: new class = rank of the canonical member */
__global__ void
k3_refine_assign_kernel(uint32_t m, const uint32_t *rep, const uint32_t *cmin, const uint64_t *rank, uint32_t *newcls)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < m) newcls[i] = (uint32_t) rank[cmin[rep[i]]];
}

/* output transitions: row of class c = classes of the canonical member's successors */
__global__ void
k3_out_trans_kernel(uint32_t m, uint32_t K, const uint32_t *flag, const uint64_t *rank, const uint32_t *cls,
	const uint32_t *ktrans, const uint32_t *korig, uint32_t *otrans, uint32_t *oorig)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m || !flag[i]) return;
	const uint32_t c = (uint32_t) rank[i];
	for (uint32_t k = 0; k < K; k++) {
		const uint32_t t = ktrans[(size_t) i * K + k];
		otrans[(size_t) c * K + k] = t == NONE32 ? NONE32 : cls[t];
	}
	oorig[c] = korig[i];
}

thread_local fsm_b200_det_stats tl_min_stats;

int
empty_result(Owner *own, fsm_b200_owned_desc *out)
{
	own->group_off.assign(1, 0);
	own->endid_off.assign(1, 0);
	out->desc.nstates = 0;
	out->desc.hasstart = 0;
	out->desc.group_off = own->group_off.data();
	out->desc.endid_off = own->endid_off.data();
	out->owner = own;
	return 0;
}

} // namespace

extern "C" int
fsm_b200_minimise_stats(struct fsm_b200_det_stats *st)
{
	if (st == nullptr) { errno = EINVAL; return -1; }
	*st = tl_min_stats;
	return 0;
}

extern "C" int
fsm_b200_minimise(const struct fsm_b200_desc *dfa, int device, struct fsm_b200_owned_desc *out)
{
	if (dfa == nullptr || out == nullptr || (dfa->reserved & ~FSM_B200_DESC_EAGER) != 0) {
		set_error("minimise: bad argument");
		errno = EINVAL;
		return -1;
	}
	memset(out, 0, sizeof *out);
	memset(&tl_min_stats, 0, sizeof tl_min_stats);
	const auto t_begin = std::chrono::steady_clock::now();
	const uint64_t launches0 = fsm_b200_launch_count(0);
	const uint32_t n = dfa->nstates;
	Owner *own = new (std::nothrow) Owner();
	if (own == nullptr) { errno = ENOMEM; return -1; }
	struct Guard { Owner *o; ~Guard() { delete o; } } guard{ own };

	/* must be a DFA (minimise.c:89-90 asserts fsm_all(fsm, fsm_isdfa)) */
	if (!dfa->hasstart || dfa->start >= n) { set_error("minimise: no start state"); errno = EINVAL; return -1; }
	const uint64_t G = dfa->group_off[n];
	if (G >= (1ull << 31)) { set_error("minimise: automaton too large"); errno = EINVAL; return -1; }
	for (uint32_t s = 0; s < n; s++) {
		if (dfa->eps_off != nullptr && dfa->eps_off[s + 1] != dfa->eps_off[s]) { set_error("minimise: not a DFA (epsilon edge)"); errno = EINVAL; return -1; }
		uint64_t seen[4] = { 0, 0, 0, 0 };
		for (uint64_t g = dfa->group_off[s]; g < dfa->group_off[s + 1]; g++) {
			if (dfa->group_to[g] >= n) { set_error("minimise: edge out of range"); errno = EINVAL; return -1; }
			for (int w = 0; w < 4; w++) {
				if (seen[w] & dfa->group_symbols[4 * g + w]) { set_error("minimise: not a DFA (ambiguous symbol)"); errno = EINVAL; return -1; }
				seen[w] |= dfa->group_symbols[4 * g + w];
			}
		}
	}

	uint8_t class_of[256], rep[256];
	const uint32_t K = G > 0 ? byte_classes(dfa, G, class_of, rep) : 1;
	if (G == 0) { rep[0] = 0; memset(class_of, 0, sizeof class_of); }

	/* initial classes: 0 for non-end states; end states by end-id set (minimise.c:733-) */
	std::vector<uint32_t> cls0(n, 0);
	const uint64_t *xoff = nullptr; const uint32_t *xids = nullptr;
	const bool has_eager = eagerhost::eh_get(dfa, &xoff, &xids);
	if (has_eager) {
		/* eager-output sets separate states too, the way the reference sees them (eager_host.h) */
		eagerhost::eh_initial_classes(dfa, xoff, xids, cls0);
	} else {
		std::map<std::vector<uint32_t>, uint32_t> ids;
		for (uint32_t s = 0; s < n; s++) {
			if (!dfa->is_end[s]) continue;
			std::vector<uint32_t> key;
			if (dfa->endid_off != nullptr) key.assign(dfa->endids + dfa->endid_off[s], dfa->endids + dfa->endid_off[s + 1]);
			auto it = ids.find(key);
			if (it == ids.end()) it = ids.emplace(std::move(key), (uint32_t) ids.size() + 1).first;
			cls0[s] = it->second;
		}
	}

	CK(cudaSetDevice(device));
	{
		cudaMemPool_t mp;
		uint64_t keep_all = UINT64_MAX;
		if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep_all);
	}
	cudaStream_t st;
	(void) cudaGetLastError();          /* CK_SYNC reports launch failures of THIS call only */
	CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamSynchronize(s); cudaStreamDestroy(s); } } sguard{ st };
	Scanner scan; scan.st = st;

	std::vector<uint32_t> h_goff(n + 1);
	for (uint32_t s = 0; s <= n; s++) h_goff[s] = (uint32_t) dfa->group_off[s];
	DBuf<uint32_t> d_goff, d_gto, d_trans, d_changed, d_keep, d_cls0;
	DBuf<uint64_t> d_gsym, d_gcls, d_newid;
	DBuf<uint8_t> d_rep, d_reach, d_co;
	const uint64_t NK = (uint64_t) n * K;
	if (d_goff.reserve(n + 1, false, st) || d_gto.reserve(G + 1, false, st) || d_gsym.reserve(4 * G + 4, false, st) ||
	    d_gcls.reserve(4 * G + 4, false, st) || d_rep.reserve(256, false, st) || d_trans.reserve(NK + 1, false, st) ||
	    d_reach.reserve(n + 1, false, st) || d_co.reserve(n + 1, false, st) || d_changed.reserve(4, false, st) ||
	    d_keep.reserve(n + 1, false, st) || d_newid.reserve(n + 2, false, st) || d_cls0.reserve(n + 1, false, st)) return -1;
	CK(cudaMemcpyAsync(d_goff.p, h_goff.data(), (n + 1) * 4, cudaMemcpyHostToDevice, st));
	if (G) {
		CK(cudaMemcpyAsync(d_gto.p, dfa->group_to, G * 4, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(d_gsym.p, dfa->group_symbols, G * 32, cudaMemcpyHostToDevice, st));
	}
	CK(cudaMemcpyAsync(d_rep.p, rep, 256, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_co.p, dfa->is_end, n, cudaMemcpyHostToDevice, st));      /* co-reach seed = end states */
	CK(cudaMemcpyAsync(d_cls0.p, cls0.data(), n * 4, cudaMemcpyHostToDevice, st));
	if (G) { k2_group_classmask_kernel<<<blocks_for(G), 256, 0, st>>>(d_gsym.p, (uint32_t) G, d_rep.p, K, d_gcls.p); count_launch(); }
	k3_build_trans_kernel<<<blocks_for(n, 128), 128, 0, st>>>(n, d_goff.p, d_gto.p, d_gcls.p, K, d_trans.p); count_launch();

	/* ---- trim ---- */
	CK(cudaMemsetAsync(d_reach.p, 0, n, st));
	{
		const uint8_t one = 1;
		CK(cudaMemcpyAsync(d_reach.p + dfa->start, &one, 1, cudaMemcpyHostToDevice, st));
	}
	for (uint32_t pass = 0; pass < 2; pass++) {
		for (uint32_t it = 0; it <= n; it++) {
			uint32_t changed = 0;
			CK(cudaMemsetAsync(d_changed.p, 0, 4, st));
			/* a few sweeps per host round trip */
			for (int rep_i = 0; rep_i < 8; rep_i++) {
				if (pass == 0) k3_reach_kernel<<<blocks_for(n), 256, 0, st>>>(n, K, d_trans.p, d_reach.p, d_changed.p);
				else k3_coreach_kernel<<<blocks_for(n), 256, 0, st>>>(n, K, d_trans.p, d_co.p, d_changed.p);
				count_launch();
			}
			CK(cudaMemcpyAsync(&changed, d_changed.p, 4, cudaMemcpyDeviceToHost, st));
			CK_SYNC(st);
			if (!changed) break;
		}
	}
	k3_keep_kernel<<<blocks_for(n), 256, 0, st>>>(n, d_reach.p, d_co.p, d_keep.p); count_launch();
	if (scan.run<uint32_t>(d_keep.p, d_newid.p, n) != 0) return -1;
	uint64_t m64 = 0;
	uint32_t start_keep = 0;
	CK(cudaMemcpyAsync(&m64, d_newid.p + n, 8, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(&start_keep, d_keep.p + dfa->start, 4, cudaMemcpyDeviceToHost, st));
	CK_SYNC(st);
	const uint32_t m = (uint32_t) m64;
	if (m == 0 || !start_keep) {           /* minimise.c:98-101: nothing can match */
		guard.o = nullptr;
		return empty_result(own, out);
	}

	/* ---- refine ---- */
	const uint64_t MK = (uint64_t) m * K;
	DBuf<uint32_t> d_ktrans, d_cls, d_newcls, d_korig, d_repst, d_cmin, d_flag;
	DBuf<uint64_t> d_rank, d_slots;
	uint64_t tab_cap = 1024;
	while (tab_cap < 2ull * m) tab_cap <<= 1;
	if (d_ktrans.reserve(MK + 1, false, st) || d_cls.reserve(m + 1, false, st) || d_newcls.reserve(m + 1, false, st) ||
	    d_korig.reserve(m + 1, false, st) || d_repst.reserve(m + 1, false, st) || d_cmin.reserve(m + 1, false, st) ||
	    d_flag.reserve(m + 2, false, st) || d_rank.reserve(m + 2, false, st) || d_slots.reserve(tab_cap, false, st)) return -1;
	k3_compact_kernel<<<blocks_for(n, 128), 128, 0, st>>>(n, K, d_trans.p, d_keep.p, d_newid.p, d_cls0.p, d_ktrans.p, d_cls.p, d_korig.p); count_launch();
	uint64_t ncls = 0, rounds = 0;
	for (;;) {
		rounds++;
		k2_fill_u64_kernel<<<blocks_for(tab_cap), 256, 0, st>>>(d_slots.p, EMPTY64, tab_cap); count_launch();
		k2_fill_u32_kernel<<<blocks_for(m), 256, 0, st>>>(d_cmin.p, NONE32, m); count_launch();
		CK(cudaMemsetAsync(d_flag.p, 0, (m + 1) * 4, st));
		k3_refine_probe_kernel<<<blocks_for(m, 128), 128, 0, st>>>(RTable{ d_slots.p, tab_cap - 1 }, m, K, d_cls.p, d_ktrans.p, d_repst.p, d_cmin.p); count_launch();
		k3_refine_flag_kernel<<<blocks_for(m), 256, 0, st>>>(m, d_repst.p, d_cmin.p, d_flag.p); count_launch();
		if (scan.run<uint32_t>(d_flag.p, d_rank.p, m) != 0) return -1;
		k3_refine_assign_kernel<<<blocks_for(m), 256, 0, st>>>(m, d_repst.p, d_cmin.p, d_rank.p, d_newcls.p); count_launch();
		uint64_t cnt = 0;
		CK(cudaMemcpyAsync(&cnt, d_rank.p + m, 8, cudaMemcpyDeviceToHost, st));
		CK_SYNC(st);
		std::swap(d_cls.p, d_newcls.p);
		std::swap(d_cls.cap, d_newcls.cap);
		if (cnt == ncls) break;              /* classes only split: same count == same partition */
		ncls = cnt;
		if (rounds > (uint64_t) m + 2) { set_error("minimise: refinement did not converge"); errno = EIO; return -1; }
	}
	/* d_flag / d_rank of the last round describe the final classes (canonical = smallest member) */

	/* ---- emit ---- */
	const uint32_t D = (uint32_t) ncls;
	DBuf<uint32_t> d_otrans, d_oorig, d_ng, d_ogto;
	DBuf<uint64_t> d_cmask, d_ogoff, d_ogsym;
	if (d_otrans.reserve((size_t) D * K + 1, false, st) || d_oorig.reserve(D + 1, false, st) || d_ng.reserve(D + 1, false, st) ||
	    d_cmask.reserve(1024, false, st) || d_ogoff.reserve(D + 2, false, st)) return -1;
	k3_out_trans_kernel<<<blocks_for(m, 128), 128, 0, st>>>(m, K, d_flag.p, d_rank.p, d_cls.p, d_ktrans.p, d_korig.p, d_otrans.p, d_oorig.p); count_launch();
	uint64_t class_mask[256][4];
	memset(class_mask, 0, sizeof class_mask);
	for (int c = 0; c < 256; c++) class_mask[class_of[c]][c >> 6] |= 1ull << (c & 63);
	CK(cudaMemcpyAsync(d_cmask.p, class_mask, sizeof class_mask, cudaMemcpyHostToDevice, st));
	k2_emit_count_kernel<<<blocks_for(D, 128), 128, 0, st>>>(d_otrans.p, D, K, d_ng.p); count_launch();
	if (scan.run<uint32_t>(d_ng.p, d_ogoff.p, D) != 0) return -1;
	own->group_off.assign(D + 1, 0);
	CK(cudaMemcpyAsync(own->group_off.data(), d_ogoff.p, (D + 1) * 8, cudaMemcpyDeviceToHost, st));
	CK_SYNC(st);
	const uint64_t NG = own->group_off[D];
	if (d_ogto.reserve(NG + 1, false, st) || d_ogsym.reserve(4 * NG + 4, false, st)) return -1;
	k2_emit_fill_kernel<<<blocks_for(D, 128), 128, 0, st>>>(d_otrans.p, D, K, d_cmask.p, d_ogoff.p, d_ogto.p, d_ogsym.p); count_launch();
	own->groups_alloc(NG);           /* pinned pool blocks for large results (common.h) */
	std::vector<uint32_t> h_oorig(D);
	uint32_t start_cls = 0;
	uint64_t start_new = 0;
	if (NG) {
		CK(cudaMemcpyAsync(own->gto(), d_ogto.p, NG * 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(own->gsym(), d_ogsym.p, NG * 32, cudaMemcpyDeviceToHost, st));
	}
	CK(cudaMemcpyAsync(h_oorig.data(), d_oorig.p, D * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(&start_new, d_newid.p + dfa->start, 8, cudaMemcpyDeviceToHost, st));
	CK_SYNC(st);
	CK(cudaMemcpyAsync(&start_cls, d_cls.p + start_new, 4, cudaMemcpyDeviceToHost, st));
	CK_SYNC(st);

	own->is_end.assign(D, 0);
	own->endid_off.assign(D + 1, 0);
	for (uint32_t c = 0; c < D; c++) {
		const uint32_t s = h_oorig[c];
		own->is_end[c] = dfa->is_end[s];
		if (dfa->is_end[s] && dfa->endid_off != nullptr) {
			own->endids.insert(own->endids.end(), dfa->endids + dfa->endid_off[s], dfa->endids + dfa->endid_off[s + 1]);
		}
		own->endid_off[c + 1] = own->endids.size();
	}
	if (has_eager) {
		/* a merged state fires the union of its members' ids (fsm_consolidate, consolidate.c:306-315) */
		/* d_cls: output state of every kept state (what k3_out_trans_kernel writes into the rows) */
		std::vector<uint32_t> h_cls(m), h_korig(m);
		CK(cudaMemcpyAsync(h_cls.data(), d_cls.p, (size_t) m * 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(h_korig.data(), d_korig.p, (size_t) m * 4, cudaMemcpyDeviceToHost, st));
		CK_SYNC(st);
		std::vector<std::vector<uint32_t>> acc(D);
		for (uint32_t i = 0; i < m; i++) {
			const uint32_t c = h_cls[i];
			if (c >= D) { set_error("minimise: class index out of range"); errno = EIO; return -1; }
			const uint32_t s = h_korig[i];
			acc[c].insert(acc[c].end(), xids + xoff[s], xids + xoff[s + 1]);
		}
		own->eager_off.assign(D + 1, 0);
		for (uint32_t c = 0; c < D; c++) {
			eagerhost::eh_sort_unique(acc[c]);
			own->eager_ids.insert(own->eager_ids.end(), acc[c].begin(), acc[c].end());
			own->eager_off[c + 1] = own->eager_ids.size();
		}
		if (own->eager_ids.empty()) own->eager_off.clear();
	}
	if (NG == 0) { own->group_to.assign(1, 0); own->group_sym.assign(4, 0); }
	if (own->endids.empty()) own->endids.push_back(0);

	out->desc.nstates = D;
	out->desc.start = start_cls;
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
	tl_min_stats.ms_total = ms_since(t_begin);
	tl_min_stats.dfa_states = D;
	tl_min_stats.dfa_groups = NG;
	tl_min_stats.rounds = rounds;
	tl_min_stats.kernel_launches = fsm_b200_launch_count(0) - launches0;
	return 0;
}
