/*
 * shim_det_bench.c -- BASELINE config 5 through libfsm's OWN API, timing fsm_determinise.
 * The same source is linked twice (libfsm_b200/shim/Makefile): against the shim
 * (build/shim/det_bench_b200: fsm_determinise -> K2 on the GPU, incl. struct fsm <-> flat
 * marshalling on both sides) and against the unmodified reference
 * (build/shim/det_bench_ref).  Prints one JSON line.
 *   det_bench [words=2000] [length=50] [seed=12345]
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <fsm/fsm.h>

static unsigned long long rng_state;
static unsigned
next_letter(void)
{
	rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
	return (unsigned) ((rng_state * 2685821657736338717ull) >> 33) % 26u;
}

static double
now_ms(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
}

int
main(int argc, char **argv)
{
	const unsigned words = argc > 1 ? (unsigned) atoi(argv[1]) : 2000;
	const unsigned length = argc > 2 ? (unsigned) atoi(argv[2]) : 50;
	struct fsm *fsm = fsm_new(NULL);
	fsm_state_t start;
	unsigned w, j;
	double t0, t1;

	rng_state = argc > 3 ? strtoull(argv[3], NULL, 10) : 12345ull;
	if (fsm == NULL || !fsm_addstate(fsm, &start)) return 2;
	fsm_setstart(fsm, start);
	if (!fsm_addedge_any(fsm, start, start)) return 2;         /* the /./ self-loop */
	for (w = 0; w < words; w++) {
		fsm_state_t prev = start;
		for (j = 0; j < length; j++) {
			fsm_state_t s;
			if (!fsm_addstate(fsm, &s)) return 2;
			if (!fsm_addedge_literal(fsm, prev, s, (char) ('a' + next_letter()))) return 2;
			prev = s;
		}
		fsm_setend(fsm, prev, 1);
		if (!fsm_endid_set(fsm, prev, w)) return 2;
	}
	{
		const unsigned nfa_states = fsm_countstates(fsm);
		/* warm-up on a throw-away copy so one-off costs (CUDA context) are not timed */
		struct fsm *tmp = fsm_new(NULL);
		fsm_state_t a, b;
		if (tmp != NULL && fsm_addstate(tmp, &a) && fsm_addstate(tmp, &b)) {
			fsm_setstart(tmp, a); fsm_addedge_literal(tmp, a, b, 'x'); fsm_setend(tmp, b, 1);
			(void) fsm_determinise(tmp);
		}
		if (tmp != NULL) fsm_free(tmp);
		t0 = now_ms();
		if (!fsm_determinise(fsm)) { fprintf(stderr, "fsm_determinise failed\n"); return 1; }
		t1 = now_ms();
		printf("{\"nfa_states\": %u, \"dfa_states\": %u, \"dfa_edges\": %u, \"determinise_ms\": %.3f}\n",
		    nfa_states, fsm_countstates(fsm), fsm_countedges(fsm), t1 - t0);
	}
	fsm_free(fsm);
	return 0;
}
