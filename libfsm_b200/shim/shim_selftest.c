/*
 * shim_selftest.c -- exercises the libfsm-facing boundary the way a libfsm user would:
 * re_comp -> fsm_determinise (K2 through the shim) -> fsm_minimise (K3 through the shim) ->
 * fsm_exec per input and fsm_exec_batch for all inputs (K1 / K1b through the shim), then
 * checks that both agree and that fsm_endid_get on `*end` returns the pattern's id.
 * Built by libfsm_b200/shim/Makefile against the reference headers; run by
 * tests/test_gpu_shim.py on the GPU box.  Exit status 0 = all checks passed.
 */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fsm/fsm.h>
#include <fsm/bool.h>
#include <re/re.h>

#include "fsm_b200_shim.h"

static int
check(int cond, const char *what)
{
	if (!cond) fprintf(stderr, "FAIL: %s (errno %d)\n", what, errno);
	return cond ? 0 : 1;
}

int
main(void)
{
	static const char *patterns[] = { "^abc[0-9]+x$", "^(GET|POST) /[a-z]+", "hello", "a[ -~]{7}\\z" };
	static const char *inputs[] = { "abc123x", "abc12", "GET /index", "POST /x y", "say hello world", "nope",
	                                "", "xxa1234567", "abcx", "abc9x" };
	enum { NP = sizeof patterns / sizeof patterns[0], NI = sizeof inputs / sizeof inputs[0] };
	struct fsm *fsms[NP], *u;
	struct re_err err;
	int fails = 0;
	size_t i;

	for (i = 0; i < NP; i++) {
		const char *s = patterns[i];
		fsms[i] = re_comp(RE_PCRE, fsm_sgetc, &s, NULL, RE_FLAGS_NONE, &err);
		fails += check(fsms[i] != NULL, "re_comp");
		fails += check(fsm_determinise(fsms[i]) == 1, "fsm_determinise (shim -> K2)");
		fails += check(fsm_minimise(fsms[i]) == 1, "fsm_minimise");
		fails += check(fsm_setendid(fsms[i], (fsm_end_id_t) (100 + i)) == 1, "fsm_setendid");
	}
	u = fsm_union_array(NP, fsms, NULL);
	fails += check(u != NULL, "fsm_union_array");
	fails += check(fsm_determinise(u) == 1, "fsm_determinise of the union (shim -> K2)");

	{
		/* an NFA must be refused exactly like the reference does: -1 / EINVAL */
		const char *s = "ab*c|abd";
		struct fsm *nfa = re_comp(RE_PCRE, fsm_sgetc, &s, NULL, RE_FLAGS_NONE, &err);
		const char *in = "abc";
		fsm_state_t end = 12345;
		errno = 0;
		fails += check(fsm_exec(nfa, fsm_sgetc, &in, &end, NULL) == -1 && errno == EINVAL && end == 12345, "NFA refused with EINVAL");
		fsm_free(nfa);
	}

	{
		unsigned char base[1024];
		uint64_t offsets[NI + 1];
		struct fsm_b200_result rec[NI];
		size_t off = 0;
		for (i = 0; i < NI; i++) {
			offsets[i] = off;
			memcpy(base + off, inputs[i], strlen(inputs[i]));
			off += strlen(inputs[i]);
		}
		offsets[NI] = off;
		fails += check(fsm_exec_batch(u, base, offsets, NI, rec) == 0, "fsm_exec_batch");
		for (i = 0; i < NI; i++) {
			const char *s = inputs[i];
			fsm_state_t end = 0;
			const int r = fsm_exec(u, fsm_sgetc, &s, &end, NULL);
			if (r != rec[i].ret || (r == 1 && end != rec[i].end)) {
				fprintf(stderr, "FAIL: input %zu \"%s\": fsm_exec %d/%u vs batch %d/%u\n", i, inputs[i], r, end, rec[i].ret, rec[i].end);
				fails++;
			}
			if (r == 1) {
				fsm_end_id_t ids[8];
				const size_t n = fsm_endid_count(u, end);
				fails += check(n >= 1 && n <= 8 && fsm_endid_get(u, end, n, ids) == 1 && ids[0] >= 100 && ids[0] < 100 + NP, "end ids of *end");
			} else {
				/* cursor left one past the byte without an edge, or at the end of the string */
				const size_t consumed = (size_t) rec[i].consumed, len = strlen(inputs[i]);
				const size_t want = consumed < len ? consumed + 1 : len;
				fails += check((size_t) (s - inputs[i]) == want, "fsm_sgetc cursor after a failed match");
			}
			printf("%-18s ret=%d end=%u consumed=%llu\n", inputs[i], rec[i].ret, rec[i].end, (unsigned long long) rec[i].consumed);
		}
	}
	fsm_free(u);
	if (fails) { fprintf(stderr, "%d check(s) failed\n", fails); return 1; }
	printf("shim selftest ok\n");
	return 0;
}
