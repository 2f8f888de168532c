/*
 * refnum_host.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Runs the product's reference-numbering functions (libfsm_b200/csrc/refnum.h, the same
 * inline code the K2 kernels call on the device) on the CPU, around a deliberately naive
 * subset construction, so that `pytest -m "not gpu"` can check them against the numbering the
 * reference recorded in tests/golden/golden_determinise.npz and against the compiled
 * reference.  Only tests load this; libfsm_b200.so never links it.
 */
#include <algorithm>
#include <array>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "../libfsm_b200/csrc/refnum.h"

extern "C" {
#include "fsm_oracle.h"
}

static uint32_t
classes_of(const struct fsm_b200_desc *d, uint8_t class_of[256])
{
	/* refine {all symbols} by every label set */
	uint32_t K = 1;
	memset(class_of, 0, 256);
	const uint64_t G = d->group_off[d->nstates];
	for (uint64_t g = 0; g < G; g++) {
		const uint64_t *sym = &d->group_symbols[4 * g];
		int split_to[256];
		for (uint32_t k = 0; k < K; k++) split_to[k] = -1;
		bool seen_out[256] = { false };
		for (int c = 0; c < 256; c++) if (!((sym[c >> 6] >> (c & 63)) & 1)) seen_out[class_of[c]] = true;
		const uint32_t K0 = K;
		for (int c = 0; c < 256; c++) {
			if (!((sym[c >> 6] >> (c & 63)) & 1)) continue;
			const uint32_t k = class_of[c];
			if (k >= K0 || !seen_out[k]) continue;       /* whole class inside the label set */
			if (split_to[k] < 0) split_to[k] = (int) K++;
			class_of[c] = (uint8_t) split_to[k];
		}
	}
	return K;
}

namespace {
struct Numbered {
	uint32_t D = 0, K = 0;
	uint8_t class_of[256];
	std::vector<uint32_t> trans;                 /* [D][K], construction numbering */
	std::vector<std::vector<uint32_t>> sets;     /* construction numbering */
	std::vector<uint32_t> perm;                  /* construction -> reference numbering */
	std::vector<uint8_t> aend;                   /* per NFA state: some closure member is an end state */
	std::vector<uint64_t> cl_off;
	std::vector<uint32_t> cl_to;
};
}

static int
numbered_determinise(const struct fsm_b200_desc *nfa, uint32_t state_limit, Numbered &N)
{
	const uint32_t n = nfa->nstates;
	uint8_t *class_of = N.class_of;

	uint64_t *cl_off = nullptr; uint32_t *cl_to = nullptr;
	if (oracle_epsilon_closure(nfa, &cl_off, &cl_to) != 0) return -1;

	const uint32_t K = classes_of(nfa, class_of);
	uint8_t rep[256];
	for (int c = 255; c >= 0; c--) rep[class_of[c]] = (uint8_t) c;

	/* (state, class) adjacency: sorted, duplicates kept */
	std::vector<std::vector<uint32_t>> lists((size_t) n * K);
	std::vector<uint8_t> aend(n, 0);
	for (uint32_t s = 0; s < n; s++) {
		for (uint64_t ci = cl_off[s]; ci < cl_off[s + 1]; ci++) {
			const uint32_t e = cl_to[ci];
			aend[s] |= nfa->is_end[e];
			for (uint64_t g = nfa->group_off[e]; g < nfa->group_off[e + 1]; g++) {
				const uint64_t *sym = &nfa->group_symbols[4 * g];
				for (uint32_t k = 0; k < K; k++) {
					const int c = rep[k];
					if ((sym[c >> 6] >> (c & 63)) & 1) lists[(size_t) s * K + k].push_back(nfa->group_to[g]);
				}
			}
		}
	}
	N.cl_off.assign(cl_off, cl_off + n + 1);
	N.cl_to.assign(cl_to, cl_to + cl_off[n]);
	free(cl_off); free(cl_to);
	std::vector<uint64_t> adj_off((size_t) n * K + 1, 0);
	std::vector<uint32_t> adj_to;
	for (size_t i = 0; i < lists.size(); i++) {
		std::sort(lists[i].begin(), lists[i].end());
		adj_to.insert(adj_to.end(), lists[i].begin(), lists[i].end());
		adj_off[i + 1] = adj_to.size();
	}
	adj_to.push_back(0);

	/* naive subset construction, discovery order over classes */
	std::map<std::vector<uint32_t>, uint32_t> ids;
	std::vector<std::vector<uint32_t>> sets;
	std::vector<uint32_t> trans;
	sets.push_back({ nfa->start });
	ids[sets[0]] = 0;
	for (uint32_t cur = 0; cur < sets.size(); cur++) {
		trans.resize((size_t) (cur + 1) * K, UINT32_MAX);
		for (uint32_t k = 0; k < K; k++) {
			std::vector<uint32_t> dst;
			for (uint32_t m : sets[cur]) {
				const std::vector<uint32_t> &l = lists[(size_t) m * K + k];
				dst.insert(dst.end(), l.begin(), l.end());
			}
			if (dst.empty()) continue;
			std::sort(dst.begin(), dst.end());
			dst.erase(std::unique(dst.begin(), dst.end()), dst.end());
			auto it = ids.find(dst);
			uint32_t id;
			if (it == ids.end()) {
				if (state_limit != 0 && sets.size() > state_limit) return 1;   /* too big for a test */
				id = (uint32_t) sets.size(); ids[dst] = id; sets.push_back(dst);
			}
			else id = it->second;
			trans[(size_t) cur * K + k] = id;
		}
	}
	const uint32_t D = (uint32_t) sets.size();

	/* the product code under test */
	std::vector<uint16_t> leaf((size_t) n * K), leaf_m(n);
	for (uint32_t s = 0; s < n; s++) leaf_m[s] = (uint16_t) rn_leaf_ranks(&adj_off[(size_t) s * K], adj_to.data(), K, &leaf[(size_t) s * K]);
	uint32_t kmax = 1;
	for (const auto &s : sets) kmax = std::max<uint32_t>(kmax, (uint32_t) s.size());
	std::vector<uint16_t> bufs((size_t) rn_depth_for(kmax) * K);
	std::vector<uint32_t> order((size_t) D * K, UINT32_MAX);
	std::vector<uint16_t> order_m(D);
	for (uint32_t d = 0; d < D; d++) {
		order_m[d] = (uint16_t) rn_state_order(sets[d].data(), (uint32_t) sets[d].size(), leaf.data(), leaf_m.data(), K,
		    bufs.data(), &trans[(size_t) d * K], &order[(size_t) d * K]);
	}
	std::vector<uint32_t> perm;
	rn_lifo_numbering(D, K, order.data(), order_m.data(), perm);
	for (uint32_t d = 0; d < D; d++) if (perm[d] >= D) { errno = EDOM; return -1; }
	N.D = D; N.K = K; N.trans.swap(trans); N.sets.swap(sets); N.perm.swap(perm); N.aend.swap(aend);
	return 0;
}

extern "C" int
refnum_host_determinise(const struct fsm_b200_desc *nfa, uint32_t state_limit, uint32_t *out_nstates, uint32_t **out_table, uint8_t **out_end)
{
	*out_nstates = 0; *out_table = nullptr; *out_end = nullptr;
	if (!nfa->hasstart || nfa->nstates == 0) return 0;
	Numbered N;
	const int rc = numbered_determinise(nfa, state_limit, N);
	if (rc != 0) return rc;
	const uint32_t D = N.D, K = N.K;
	const uint8_t *class_of = N.class_of;
	const std::vector<uint32_t> &trans = N.trans, &perm = N.perm;
	const std::vector<std::vector<uint32_t>> &sets = N.sets;
	const std::vector<uint8_t> &aend = N.aend;

	uint32_t *table = (uint32_t *) malloc((size_t) D * 256 * sizeof *table);
	uint8_t *end = (uint8_t *) calloc(D, 1);
	if (table == nullptr || end == nullptr) { free(table); free(end); errno = ENOMEM; return -1; }
	for (uint32_t d = 0; d < D; d++) {
		uint32_t *row = table + (size_t) perm[d] * 256;
		for (int c = 0; c < 256; c++) {
			const uint32_t t = trans[(size_t) d * K + class_of[c]];
			row[c] = t == UINT32_MAX ? UINT32_MAX : perm[t];
		}
		for (uint32_t m : sets[d]) end[perm[d]] |= aend[m];
	}
	*out_nstates = D; *out_table = table; *out_end = end;
	return 0;
}

/* The same, as a complete flat description in the reference's numbering (one group per
 * destination, ascending; end bits and end-id sets carried as determinise.c:236-266 does):
 * what fsm_b200_determinise_ex(FSM_B200_DET_REFERENCE_NUMBERING) returns.  Used by the stub
 * engine so that the shim's struct fsm rebuild can be compared with the reference's fsm(1)
 * output TEXTUALLY on the CPU.  Returns 0, 1 (state limit) or -1. */
extern "C" int
refnum_host_determinise_desc(const struct fsm_b200_desc *nfa, uint32_t state_limit, struct oracle_owned_desc *out)
{
	memset(out, 0, sizeof *out);
	if (!nfa->hasstart || nfa->nstates == 0) {
		uint64_t *z = (uint64_t *) calloc(2, sizeof *z), *z2 = (uint64_t *) calloc(2, sizeof *z2);
		out->desc.group_off = z; out->desc.endid_off = z2; out->blocks[0] = z; out->blocks[1] = z2;
		return 0;
	}
	Numbered N;
	const int rc = numbered_determinise(nfa, state_limit, N);
	if (rc != 0) return rc;
	const uint32_t D = N.D, K = N.K;
	std::vector<uint32_t> inv(D);
	for (uint32_t d = 0; d < D; d++) inv[N.perm[d]] = d;

	std::vector<uint64_t> goff(D + 1, 0), gsym, ioff(D + 1, 0);
	std::vector<uint32_t> gto, ids;
	std::vector<uint8_t> end(D, 0);
	for (uint32_t s = 0; s < D; s++) {
		const uint32_t d = inv[s];
		std::map<uint32_t, std::array<uint64_t, 4>> groups;
		for (int c = 0; c < 256; c++) {
			const uint32_t t = N.trans[(size_t) d * K + N.class_of[c]];
			if (t == UINT32_MAX) continue;
			auto &m = groups[N.perm[t]];
			m[c >> 6] |= 1ull << (c & 63);
		}
		for (auto &g : groups) { gto.push_back(g.first); gsym.insert(gsym.end(), g.second.begin(), g.second.end()); }
		goff[s + 1] = gto.size();
		std::vector<uint32_t> my;
		for (uint32_t m : N.sets[d]) {
			if (!N.aend[m]) continue;
			end[s] = 1;
			if (nfa->endid_off == nullptr) continue;
			for (uint64_t ci = N.cl_off[m]; ci < N.cl_off[m + 1]; ci++) {
				const uint32_t e = N.cl_to[ci];
				if (!nfa->is_end[e]) continue;
				for (uint64_t q = nfa->endid_off[e]; q < nfa->endid_off[e + 1]; q++) my.push_back(nfa->endids[q]);
			}
		}
		std::sort(my.begin(), my.end());
		my.erase(std::unique(my.begin(), my.end()), my.end());
		ids.insert(ids.end(), my.begin(), my.end());
		ioff[s + 1] = ids.size();
	}
	auto dup = [](const void *p, size_t bytes) { void *q = malloc(bytes ? bytes : 8); if (q && bytes) memcpy(q, p, bytes); return q; };
	out->desc.nstates = D; out->desc.start = 0; out->desc.hasstart = 1;
	out->desc.is_end = (const uint8_t *) (out->blocks[0] = dup(end.data(), D));
	out->desc.group_off = (const uint64_t *) (out->blocks[1] = dup(goff.data(), (D + 1) * 8));
	out->desc.group_symbols = (const uint64_t *) (out->blocks[2] = dup(gsym.data(), gsym.size() * 8));
	out->desc.group_to = (const uint32_t *) (out->blocks[3] = dup(gto.data(), gto.size() * 4));
	out->desc.endid_off = (const uint64_t *) (out->blocks[4] = dup(ioff.data(), (D + 1) * 8));
	out->desc.endids = (const uint32_t *) (out->blocks[5] = dup(ids.data(), ids.size() * 4));
	return 0;
}
