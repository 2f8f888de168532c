/* k2_determinise.cu -- K2: subset construction (placeholder until the kernels land). */
#include <cstring>
#include "common.h"
using namespace fsmb200;

extern "C" int
fsm_b200_determinise(const struct fsm_b200_desc *, int, size_t, struct fsm_b200_owned_desc *)
{ set_error("determinise: not implemented yet"); errno = ENOTSUP; return -1; }

extern "C" void
fsm_b200_desc_free(struct fsm_b200_owned_desc *) {}

extern "C" int
fsm_b200_determinise_stats(struct fsm_b200_det_stats *st)
{ if (st) memset(st, 0, sizeof *st); return 0; }
