/*
 * eager_host.h -- host-side pieces of eager-output support (include/fsm/fsm.h:273-336), shared by
 * dfa_compile.cu, k2_determinise.cu and k3_minimise.cu.  Plain C++ with no CUDA in it, so that
 * oracle/eager_host_test.cpp can compile the SAME functions with g++ and check them against the
 * oracle and the compiled reference on a machine without a GPU.
 *
 * What the reference does with eager outputs on the path this engine replaces:
 *   exec         a state's ids fire every time it is entered, the start state included
 *                (src/libfsm/exec.c:126-130,140-144)                       -> eh_build_masks
 *   determinise  epsilon removal copies the ids of every state of s's closure onto s
 *                (epsilons.c:221-253); a DFA state then gets the ids of all its members
 *                (determinise.c:2614-2636)                                 -> eh_union_over
 *   minimise     states are told apart by their eager-id set as well (same_end_metadata,
 *                minimise.c:705-731) -- but split_ecs_by_end_metadata only LOOKS at a state's ids
 *                while it walks its initial class (states at equal shortest distance from an end
 *                state, in descending state order) and stops at the first state that is neither
 *                an end state nor has eager outputs (minimise.c:771-782): ids behind that state are
 *                not seen.  Merged states get the union (consolidate.c:306-315)
 *                                                                          -> eh_initial_classes
 */
#ifndef FSM_B200_EAGER_HOST_H
#define FSM_B200_EAGER_HOST_H

#include <stdint.h>

#include <algorithm>
#include <map>
#include <utility>
#include <vector>

#include "../../include/fsm_b200.h"

namespace eagerhost {

/* eager CSR of a description, or false when it has none */
inline bool
eh_get(const struct fsm_b200_desc *d, const uint64_t **off, const uint32_t **ids)
{
	*off = nullptr; *ids = nullptr;
	if (!(d->reserved & FSM_B200_DESC_EAGER)) return false;
	const struct fsm_b200_desc_ext *x = reinterpret_cast<const struct fsm_b200_desc_ext *>(d);
	if (x->eager_off == nullptr || x->eager_off[d->nstates] == 0) return false;
	*off = x->eager_off; *ids = x->eager_ids;
	return true;
}

/* distinct ids of the whole automaton, ascending: bit b of a mask <=> id_of_bit[b] */
inline void
eh_id_list(uint32_t nstates, const uint64_t *off, const uint32_t *ids, std::vector<uint32_t> &id_of_bit)
{
	id_of_bit.assign(ids, ids + off[nstates]);
	std::sort(id_of_bit.begin(), id_of_bit.end());
	id_of_bit.erase(std::unique(id_of_bit.begin(), id_of_bit.end()), id_of_bit.end());
}

/* per-row bit masks [nrows][words]; rows >= nstates (the dead row) stay empty */
inline void
eh_build_masks(uint32_t nstates, uint32_t nrows, const uint64_t *off, const uint32_t *ids,
	const std::vector<uint32_t> &id_of_bit, uint32_t words, std::vector<uint64_t> &masks)
{
	masks.assign((size_t) nrows * words, 0);
	for (uint32_t s = 0; s < nstates; s++) {
		for (uint64_t q = off[s]; q < off[s + 1]; q++) {
			const size_t b = (size_t) (std::lower_bound(id_of_bit.begin(), id_of_bit.end(), ids[q]) - id_of_bit.begin());
			masks[(size_t) s * words + (b >> 6)] |= 1ull << (b & 63);
		}
	}
}

/* append the ids of states [first, last) (any iterator over state numbers) to acc */
template <typename It>
inline void
eh_union_over(It first, It last, const uint64_t *off, const uint32_t *ids, std::vector<uint32_t> &acc)
{
	for (; first != last; ++first) acc.insert(acc.end(), ids + off[*first], ids + off[*first + 1]);
}

inline void
eh_sort_unique(std::vector<uint32_t> &v)
{
	std::sort(v.begin(), v.end());
	v.erase(std::unique(v.begin(), v.end()), v.end());
}

/* Initial classes for the partition refinement of a DFA with eager outputs: 0 for the states
 * the reference treats as plain, otherwise one id per distinct (end bit, end-id set, SEEN eager-id
 * set).  Trimmed-away states (not reachable from the start, or unable to reach an end state:
 * fsm_trim, minimise.c:93-96) get 0; they take no part in the refinement. */
inline void
eh_initial_classes(const struct fsm_b200_desc *d, const uint64_t *xoff, const uint32_t *xids, std::vector<uint32_t> &cls0)
{
	const uint32_t n = d->nstates;
	const uint32_t NONE = 0xFFFFFFFFu;
	cls0.assign(n, 0);

	/* forward reachability, and reverse adjacency for the backward distances */
	std::vector<uint8_t> reach(n, 0);
	std::vector<uint32_t> stack, rdeg(n + 1, 0), radj;
	reach[d->start] = 1; stack.push_back(d->start);
	while (!stack.empty()) {
		const uint32_t s = stack.back(); stack.pop_back();
		for (uint64_t g = d->group_off[s]; g < d->group_off[s + 1]; g++) {
			const uint32_t t = d->group_to[g];
			if (!reach[t]) { reach[t] = 1; stack.push_back(t); }
		}
	}
	for (uint32_t s = 0; s < n; s++) for (uint64_t g = d->group_off[s]; g < d->group_off[s + 1]; g++) rdeg[d->group_to[g] + 1]++;
	for (uint32_t s = 0; s < n; s++) rdeg[s + 1] += rdeg[s];
	radj.resize(rdeg[n]);
	{
		std::vector<uint32_t> cur(rdeg.begin(), rdeg.end() - 1);
		for (uint32_t s = 0; s < n; s++) for (uint64_t g = d->group_off[s]; g < d->group_off[s + 1]; g++) radj[cur[d->group_to[g]]++] = s;
	}
	/* shortest distance to an end state over the states fsm_trim keeps (level-synchronous BFS
	 * from the reachable end states, walking only reachable predecessors) */
	std::vector<uint32_t> dist(n, NONE), frontier, next;
	for (uint32_t s = 0; s < n; s++) if (reach[s] && d->is_end[s]) { dist[s] = 0; frontier.push_back(s); }
	for (uint32_t level = 1; !frontier.empty(); level++) {
		next.clear();
		for (uint32_t t : frontier) {
			for (uint32_t k = rdeg[t]; k < rdeg[t + 1]; k++) {
				const uint32_t s = radj[k];
				if (reach[s] && dist[s] == NONE) { dist[s] = level; next.push_back(s); }
			}
		}
		frontier.swap(next);
	}
	/* which states have their eager ids looked at: per distance class, descending state order,
	 * until the first plain state */
	std::vector<uint8_t> seen(n, 0);
	{
		std::vector<uint8_t> broken;
		for (uint32_t s = n; s-- > 0; ) {
			if (dist[s] == NONE) continue;
			if (dist[s] >= broken.size()) broken.resize((size_t) dist[s] + 1, 0);
			if (broken[dist[s]]) continue;
			if (d->is_end[s] || xoff[s + 1] > xoff[s]) seen[s] = 1;
			else broken[dist[s]] = 1;
		}
	}
	std::map<std::pair<std::vector<uint32_t>, std::vector<uint32_t>>, uint32_t> end_keys, plain_keys;
	for (uint32_t s = 0; s < n; s++) {
		if (dist[s] == NONE) continue;
		std::vector<uint32_t> eids, xs;
		if (d->is_end[s] && d->endid_off != nullptr) eids.assign(d->endids + d->endid_off[s], d->endids + d->endid_off[s + 1]);
		if (seen[s]) xs.assign(xids + xoff[s], xids + xoff[s + 1]);
		if (!d->is_end[s] && xs.empty()) continue;                 /* plain: class 0 */
		auto &keys = d->is_end[s] ? end_keys : plain_keys;
		auto key = std::make_pair(std::move(eids), std::move(xs));
		auto it = keys.find(key);
		if (it == keys.end()) it = keys.emplace(std::move(key), (uint32_t) (end_keys.size() + plain_keys.size()) + 1).first;
		cls0[s] = it->second;
	}
}

} // namespace eagerhost

#endif /* FSM_B200_EAGER_HOST_H */
