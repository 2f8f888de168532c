/*
 * shim_flatten_dump.c -- host-only check of the shim's struct fsm -> flat description walk
 * (fsm_b200_flatten): compiles a PCRE with the reference's re_comp (optionally determinises with
 * the REFERENCE's code: this program is linked to the unmodified reference library plus
 * fsm_b200_shim.c's flattener only), flattens it and prints the arrays as text.  No GPU involved.
 *   flatten_dump <regex> [d]      d = fsm_determinise + fsm_minimise first
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fsm/fsm.h>
#include <re/re.h>

#include "fsm_b200_shim.h"

int
main(int argc, char **argv)
{
	struct re_err err;
	struct fsm *fsm;
	struct fsm_b200_flat flat;
	const char *s;
	uint32_t i;
	uint64_t g;

	if (argc < 2) return 2;
	s = argv[1];
	fsm = re_comp(RE_PCRE, fsm_sgetc, &s, NULL, RE_FLAGS_NONE, &err);
	if (fsm == NULL) return 3;
	if (argc > 2 && argv[2][0] == 'd') {
		if (!fsm_determinise(fsm) || !fsm_minimise(fsm)) return 4;
		fsm_setendid(fsm, 7);
	}
	if (fsm_b200_flatten(fsm, &flat) != 0) return 5;
	printf("nstates %u start %u hasstart %u\n", flat.desc.nstates, flat.desc.start, flat.desc.hasstart);
	printf("is_end");
	for (i = 0; i < flat.desc.nstates; i++) printf(" %u", flat.desc.is_end[i]);
	printf("\ngroup_off");
	for (i = 0; i <= flat.desc.nstates; i++) printf(" %llu", (unsigned long long) flat.desc.group_off[i]);
	printf("\ngroups");
	for (g = 0; g < flat.desc.group_off[flat.desc.nstates]; g++) {
		printf(" %u:%llx:%llx:%llx:%llx", flat.desc.group_to[g],
		    (unsigned long long) flat.desc.group_symbols[4 * g], (unsigned long long) flat.desc.group_symbols[4 * g + 1],
		    (unsigned long long) flat.desc.group_symbols[4 * g + 2], (unsigned long long) flat.desc.group_symbols[4 * g + 3]);
	}
	printf("\neps_off");
	for (i = 0; i <= flat.desc.nstates; i++) printf(" %llu", (unsigned long long) flat.desc.eps_off[i]);
	printf("\neps_to");
	for (g = 0; g < flat.desc.eps_off[flat.desc.nstates]; g++) printf(" %u", flat.desc.eps_to[g]);
	printf("\nendid_off");
	for (i = 0; i <= flat.desc.nstates; i++) printf(" %llu", (unsigned long long) flat.desc.endid_off[i]);
	printf("\nendids");
	for (g = 0; g < flat.desc.endid_off[flat.desc.nstates]; g++) printf(" %u", flat.desc.endids[g]);
	printf("\n");
	fsm_b200_flat_free(&flat);
	fsm_free(fsm);
	return 0;
}
