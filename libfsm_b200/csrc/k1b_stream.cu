/* k1b_stream.cu -- K1b: one long input (placeholder until the chunk-map kernels land). */
#include "common.h"
using namespace fsmb200;

extern "C" int
fsm_b200_exec_stream_host(const fsm_b200_dfa *, const uint8_t *, uint64_t, struct fsm_b200_result *)
{ set_error("exec_stream_host: not implemented yet"); errno = ENOTSUP; return -1; }

extern "C" int
fsm_b200_exec_stream_dev(const fsm_b200_dfa *, const uint8_t *, uint64_t, struct fsm_b200_result *, void *)
{ set_error("exec_stream_dev: not implemented yet"); errno = ENOTSUP; return -1; }

extern "C" int
fsm_b200_exec_stream_map_dev(const fsm_b200_dfa *, const uint8_t *, uint64_t, uint32_t *, uint64_t *, uint32_t *, void *)
{ set_error("exec_stream_map_dev: not implemented yet"); errno = ENOTSUP; return -1; }
