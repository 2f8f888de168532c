/*
 * shim_eager_selftest.c -- eager outputs through the libfsm-facing boundary, the way a libfsm
 * user reaches them: re_comp(RE_SAVE_LINKAGE_INFO) x N -> fsm_union_repeated_pattern_group
 * (the reference's own producer of eager outputs, include/fsm/bool.h) -> fsm_determinise ->
 * fsm_minimise (both through the shim: the engine carries the ids) -> per input
 *   fsm_exec + fsm_eager_output_cb   (the reference's interface, exec.c:126-144), and
 *   fsm_exec_batch_eager             (the additive batch form: one bitset per input).
 * Both must report the same set of ids per input, and the sets must be the expected ones
 * (pattern i fires id i+1 iff the input contains a match of pattern i).
 * Exit status 0 = all checks passed.
 */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fsm/fsm.h>
#include <fsm/bool.h>
#include <re/re.h>

#include "fsm_b200_shim.h"

struct seen { unsigned ids[64]; size_t n; };

static void
collect(fsm_output_id_t id, void *opaque)
{
	struct seen *s = opaque;
	size_t i;
	for (i = 0; i < s->n; i++) if (s->ids[i] == id) return;
	if (s->n < 64) s->ids[s->n++] = id;
}

static int
cmp_unsigned(const void *a, const void *b)
{
	const unsigned x = *(const unsigned *) a, y = *(const unsigned *) b;
	return x < y ? -1 : x > y;
}

int
main(void)
{
	static const char *patterns[] = { "abc", "b+c", "xyz", "q" };       /* unanchored: reported as eager outputs */
	static const struct { const char *in; unsigned want[5]; } cases[] = {
		{ "abc",     { 1, 2, 0 } },
		{ "zabcq",   { 1, 2, 4, 0 } },
		{ "xyzbbc",  { 2, 3, 0 } },
		{ "xy",      { 0 } },
		{ "",        { 0 } },
		{ "qq",      { 4, 0 } },
	};
	enum { NP = sizeof patterns / sizeof patterns[0], NC = sizeof cases / sizeof cases[0] };
	struct fsm *fsms[NP], *u;
	unsigned char base[256];
	uint64_t offsets[NC + 1], masks[NC * (FSM_B200_EAGER_MAX_IDS / 64)];
	struct fsm_b200_result rec[NC];
	uint32_t id_of_bit[FSM_B200_EAGER_MAX_IDS];
	uint32_t nbits = 0;
	size_t i, off = 0, words;
	int fails = 0;

	for (i = 0; i < NP; i++) {
		const char *s = patterns[i];
		fsms[i] = re_comp(RE_PCRE, fsm_sgetc, &s, NULL, RE_SAVE_LINKAGE_INFO, NULL);
		if (fsms[i] == NULL) { fprintf(stderr, "FAIL: re_comp %s\n", patterns[i]); return 1; }
	}
	u = fsm_union_repeated_pattern_group(NP, fsms, NULL, 1);
	if (u == NULL || fsm_determinise(u) != 1 || fsm_minimise(u) != 1) {
		fprintf(stderr, "FAIL: building the automaton (errno %d)\n", errno);
		return 1;
	}
	for (i = 0; i < NC; i++) {
		offsets[i] = off;
		memcpy(base + off, cases[i].in, strlen(cases[i].in));
		off += strlen(cases[i].in);
	}
	offsets[NC] = off;
	memset(masks, 0, sizeof masks);
	if (fsm_exec_batch_eager(u, base, offsets, NC, rec, masks, &nbits, id_of_bit) != 0) {
		fprintf(stderr, "FAIL: fsm_exec_batch_eager (errno %d)\n", errno);
		return 1;
	}
	words = (nbits + 63) / 64;
	if (nbits == 0 || nbits > NP) { fprintf(stderr, "FAIL: %u eager ids\n", nbits); fails++; }

	for (i = 0; i < NC; i++) {
		struct seen cb = { { 0 }, 0 };
		unsigned batch[64], want[8];
		size_t nb = 0, nw = 0, k;
		const char *s = cases[i].in;
		fsm_state_t end = 0;
		int r, ok;

		fsm_eager_output_set_cb(u, collect, &cb);
		r = fsm_exec(u, fsm_sgetc, &s, &end, NULL);
		fsm_eager_output_set_cb(u, NULL, NULL);
		qsort(cb.ids, cb.n, sizeof cb.ids[0], cmp_unsigned);
		for (k = 0; k < nbits; k++) if ((masks[i * words + (k >> 6)] >> (k & 63)) & 1u) batch[nb++] = id_of_bit[k];
		while (cases[i].want[nw] != 0) { want[nw] = cases[i].want[nw]; nw++; }

		ok = r == rec[i].ret && nb == cb.n && memcmp(batch, cb.ids, nb * sizeof batch[0]) == 0
		  && nb == nw && memcmp(batch, want, nw * sizeof want[0]) == 0;
		printf("%-8s ret=%d ids:", cases[i].in, rec[i].ret);
		for (k = 0; k < nb; k++) printf(" %u", batch[k]);
		printf("%s\n", ok ? "" : "   <-- MISMATCH");
		if (!ok) {
			fprintf(stderr, "FAIL: input \"%s\": fsm_exec %d with %zu ids, batch %d with %zu ids, expected %zu ids\n",
			    cases[i].in, r, cb.n, rec[i].ret, nb, nw);
			fails++;
		}
	}
	fsm_free(u);
	if (fails) { fprintf(stderr, "%d check(s) failed\n", fails); return 1; }
	printf("shim eager selftest ok\n");
	return 0;
}
