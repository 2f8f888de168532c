/*
 * shim_threads.c -- the shim's device-table cache under concurrent use.
 *
 * The reference library is unsynchronised, but fsm_exec on a const fsm is re-entrant and lx(1)
 * drives libfsm from a pthread pool (src/lx/main.c:338-475), so the side cache the shim adds
 * (struct fsm* -> compiled device table, CACHE_SLOTS entries, LRU) must be safe when several
 * threads execute DIFFERENT automata at once and force each other's entries out.
 * Together the threads own far more automata than the cache has slots (with more threads than
 * slots an entry can be evicted WHILE another thread is still executing on it) and each runs
 * fsm_exec over its own round-robin, checking each verdict against the one computed single-threaded beforehand.
 * Second phase: every thread runs fsm_exec on the SAME automaton at once, each over its
 * own 64 KiB input (so the chunked stream path with its per-DFA staging buffer is used): threads
 * with an even number expect a match, odd ones none -- a thread that saw another thread's bytes or
 * verdict fails.
 * Exit status 0 = all verdicts identical.
 */
#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fsm/fsm.h>
#include <re/re.h>

#include "fsm_b200_shim.h"

enum { MAXTHREADS = 64, NFSM = 6, NIN = 6 };
static int NTHREADS = 4, ROUNDS = 40;    /* argv[1], argv[2] */

static const char *inputs[NIN] = { "abc0x", "k7", "zzz", "", "q3q3", "m1x" };

struct job {
	struct fsm *fsm[NFSM];
	int want_ret[NFSM][NIN];
	fsm_state_t want_end[NFSM][NIN];
	int fails;
};

static void *
worker(void *arg)
{
	struct job *j = arg;
	int r, f, i;
	for (r = 0; r < ROUNDS; r++) {
		for (f = 0; f < NFSM; f++) {
			for (i = 0; i < NIN; i++) {
				const char *s = inputs[i];
				fsm_state_t end = 0;
				const int ret = fsm_exec(j->fsm[f], fsm_sgetc, &s, &end, NULL);
				if (ret != j->want_ret[f][i] || (ret == 1 && end != j->want_end[f][i])) j->fails++;
			}
		}
	}
	return NULL;
}

enum { SHARED_LEN = 64 * 1024 };
static struct fsm *shared_fsm;
struct sjob { int id, fails; char *text; };

static void *
shared_worker(void *arg)
{
	struct sjob *j = arg;
	int r;
	for (r = 0; r < ROUNDS; r++) {
		const char *s = j->text;
		fsm_state_t end = 0;
		const int ret = fsm_exec(shared_fsm, fsm_sgetc, &s, &end, NULL);
		if (ret != (j->id % 2 == 0 ? 1 : 0)) j->fails++;
		if (s != j->text + SHARED_LEN) j->fails++;          /* the cursor ends on the NUL */
	}
	return NULL;
}

static int
shared_phase(void)
{
	static struct sjob sj[MAXTHREADS];
	pthread_t th[MAXTHREADS];
	const char *pat = "[0-9]+\\.[0-9]+x";
	int t, fails = 0;
	shared_fsm = re_comp(RE_PCRE, fsm_sgetc, &pat, NULL, RE_FLAGS_NONE, NULL);
	if (shared_fsm == NULL || !fsm_determinise(shared_fsm) || !fsm_minimise(shared_fsm)) return 1;
	for (t = 0; t < NTHREADS; t++) {
		int k;
		sj[t].id = t; sj[t].fails = 0;
		sj[t].text = malloc(SHARED_LEN + 1);
		if (sj[t].text == NULL) return 1;
		for (k = 0; k < SHARED_LEN; k++) sj[t].text[k] = (char) ('0' + (k * 7 + t) % 10);
		if (t % 2 == 0) { sj[t].text[SHARED_LEN / 2 + 97 * t] = '.'; sj[t].text[SHARED_LEN / 2 + 97 * t + 3] = 'x'; }
		sj[t].text[SHARED_LEN] = '\0';
	}
	for (t = 0; t < NTHREADS; t++) pthread_create(&th[t], NULL, shared_worker, &sj[t]);
	for (t = 0; t < NTHREADS; t++) { pthread_join(th[t], NULL); fails += sj[t].fails; free(sj[t].text); }
	fsm_free(shared_fsm);
	if (fails) fprintf(stderr, "%d verdict(s) differ when %d threads share one automaton\n", fails, NTHREADS);
	return fails != 0;
}

int
main(int argc, char **argv)
{
	static struct job jobs[MAXTHREADS];
	pthread_t th[MAXTHREADS];
	int t, f, i, fails = 0;

	if (argc > 1) NTHREADS = atoi(argv[1]);
	if (argc > 2) ROUNDS = atoi(argv[2]);
	if (NTHREADS < 1 || NTHREADS > MAXTHREADS || ROUNDS < 1) { fprintf(stderr, "usage: shim_threads [threads<=64] [rounds]\n"); return 2; }

	for (t = 0; t < NTHREADS; t++) {
		for (f = 0; f < NFSM; f++) {
			char pat[64];
			const char *s = pat;
			struct re_err err;
			/* distinct automata: letter and digit depend on (thread, index) */
			snprintf(pat, sizeof pat, "^(%c[0-9]|abc%d|%c+)x?$", 'a' + (t * NFSM + f) % 26, (t + f) % 10, 'z' - f % 3);
			jobs[t].fsm[f] = re_comp(RE_PCRE, fsm_sgetc, &s, NULL, RE_FLAGS_NONE, &err);
			if (jobs[t].fsm[f] == NULL || !fsm_determinise(jobs[t].fsm[f]) || !fsm_minimise(jobs[t].fsm[f])) {
				fprintf(stderr, "FAIL: building automaton %d/%d (errno %d)\n", t, f, errno);
				return 1;
			}
			for (i = 0; i < NIN; i++) {
				const char *in = inputs[i];
				jobs[t].want_ret[f][i] = fsm_exec(jobs[t].fsm[f], fsm_sgetc, &in, &jobs[t].want_end[f][i], NULL);
				if (jobs[t].want_ret[f][i] < 0) { fprintf(stderr, "FAIL: fsm_exec errno %d\n", errno); return 1; }
			}
		}
	}
	for (t = 0; t < NTHREADS; t++) pthread_create(&th[t], NULL, worker, &jobs[t]);
	for (t = 0; t < NTHREADS; t++) { pthread_join(th[t], NULL); fails += jobs[t].fails; }
	for (t = 0; t < NTHREADS; t++) for (f = 0; f < NFSM; f++) fsm_free(jobs[t].fsm[f]);
	if (fails) { fprintf(stderr, "%d verdict(s) differ under concurrency\n", fails); return 1; }
	if (shared_phase() != 0) return 1;
	printf("shim threads ok (%d threads x %d automata x %d rounds)\n", NTHREADS, NFSM, ROUNDS);
	return 0;
}
