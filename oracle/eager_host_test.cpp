/*
 * eager_host_test.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Compiles the product's host-side eager-output functions (libfsm_b200/csrc/eager_host.h, the
 * same inline code dfa_compile.cu / k2_determinise.cu / k3_minimise.cu call) for the CPU so that
 * `pytest -m "not gpu"` can check them against the oracle and the compiled reference.
 */
#include <cerrno>
#include <cstdlib>
#include <cstring>

#include "../libfsm_b200/csrc/eager_host.h"

extern "C" {
#include "fsm_oracle.h"
}

/* fsm_minimise with the initial partition taken from eh_initial_classes */
extern "C" int
eager_host_minimise(const struct fsm_b200_desc *dfa, struct oracle_owned_desc *out)
{
	const uint64_t *xoff; const uint32_t *xids;
	if (!eagerhost::eh_get(dfa, &xoff, &xids)) return oracle_minimise(dfa, out);
	std::vector<uint32_t> cls0;
	eagerhost::eh_initial_classes(dfa, xoff, xids, cls0);
	return oracle_minimise_from_classes(dfa, cls0.data(), out);
}

/* per-state masks + id list, as dfa_compile.cu builds them: returns words, fills *nbits;
 * masks and ids are malloc'd */
extern "C" int
eager_host_masks(const struct fsm_b200_desc *dfa, uint32_t nrows, uint32_t *nbits, uint32_t **id_of_bit, uint64_t **masks)
{
	const uint64_t *xoff; const uint32_t *xids;
	*nbits = 0; *id_of_bit = nullptr; *masks = nullptr;
	if (!eagerhost::eh_get(dfa, &xoff, &xids)) return 0;
	std::vector<uint32_t> ids;
	std::vector<uint64_t> m;
	eagerhost::eh_id_list(dfa->nstates, xoff, xids, ids);
	const uint32_t words = (uint32_t) ((ids.size() + 63) / 64);
	eagerhost::eh_build_masks(dfa->nstates, nrows, xoff, xids, ids, words, m);
	*nbits = (uint32_t) ids.size();
	*id_of_bit = (uint32_t *) malloc(ids.size() * 4 + 4);
	*masks = (uint64_t *) malloc(m.size() * 8 + 8);
	memcpy(*id_of_bit, ids.data(), ids.size() * 4);
	memcpy(*masks, m.data(), m.size() * 8);
	return (int) words;
}
