/*
 * dfa_compile.cu -- flat description -> validated, dense, device-resident DFA.
 *
 * Replaces, once per DFA instead of once per fsm_exec call, the reference's
 *   fsm_all(fsm, fsm_isdfa)            src/libfsm/exec.c:106, walk/all.c:15-31,
 *                                      pred/isdfa.c:25-55, src/adt/edgeset.c:514-562
 *   fsm_getstart                       src/libfsm/exec.c:111-114, start.c:34-48
 * and turns the per-state edge groups (src/adt/edgeset.c:34-41) into the dense
 * [state][256] table the reference never materialises (print/ir.h:98-100: IR_TABLE is
 * "not yet implemented").  Lookup semantics = edge_set_find (edgeset.c:394-418).
 */
#include <cstring>
#include <new>
#include <vector>

#include "common.h"
#include "eager_host.h"

using namespace fsmb200;

namespace {

/* fsm_isdfa for every state + fsm_getstart. */
bool
desc_is_dfa(const fsm_b200_desc *d)
{
	if (!d->hasstart || d->start >= d->nstates) {
		return false;
	}
	for (uint32_t s = 0; s < d->nstates; s++) {
		if (d->eps_off != nullptr && d->eps_off[s + 1] != d->eps_off[s]) {
			return false;               /* pred/isdfa.c:43-45 */
		}
		uint64_t seen[4] = { 0, 0, 0, 0 };
		for (uint64_t g = d->group_off[s]; g < d->group_off[s + 1]; g++) {
			const uint64_t *sym = d->group_symbols + 4 * g;
			for (int w = 0; w < 4; w++) {
				if (seen[w] & sym[w]) {
					return false;       /* edgeset.c:544-551 */
				}
				seen[w] |= sym[w];
			}
		}
	}
	return true;
}

/* Two byte ranges whose 2-bit cell code ([b in R0] + 2 [b in R1]) determines the byte class.
 * Ranges start and end at class-run boundaries and stay inside one half of the byte space
 * (the kernel compares 7-bit values and selects the half with bit 7).  Returns 0 (none),
 * 1 (both ranges below 0x80) or 2 (some range at or above 0x80); an unused second range comes
 * back empty (lo > hi). */
uint32_t
find_cell_ranges(const uint8_t (&cls)[256], uint8_t (&rlo)[2], uint8_t (&rhi)[2])
{
	std::vector<int> starts, ends;          /* run boundaries, split at 0x80 */
	for (int c = 0; c < 256; c++) {
		if (c == 0 || c == 0x80 || cls[c] != cls[c - 1]) starts.push_back(c);
		if (c == 255 || c == 0x7F || cls[c] != cls[c + 1]) ends.push_back(c);
	}
	struct Rg { int lo, hi; };
	std::vector<Rg> cand;
	cand.push_back({ 1, 0 });                /* the empty range: one range may be enough */
	for (int lo : starts) {
		for (int hi : ends) {
			if (hi < lo || (lo < 0x80) != (hi < 0x80)) continue;
			cand.push_back({ lo, hi });
		}
	}
	int nclasses = 0;
	for (int c = 0; c < 256; c++) if (cls[c] + 1 > nclasses) nclasses = cls[c] + 1;
	/* cells refine classes; *cells = how many of the 4 cell codes occur */
	auto ok = [&](const Rg &a, const Rg &b, int *cells) {
		int cls_of_cell[4] = { -1, -1, -1, -1 };
		*cells = 0;
		for (int c = 0; c < 256; c++) {
			const int cell = ((c >= a.lo && c <= a.hi) ? 1 : 0) | ((c >= b.lo && c <= b.hi) ? 2 : 0);
			if (cls_of_cell[cell] < 0) { cls_of_cell[cell] = cls[c]; (*cells)++; }
			else if (cls_of_cell[cell] != cls[c]) return false;
		}
		return true;
	};
	/* prefer one cell per class (inputs of one class then share their tuple index, hence their table
	 * word: a broadcast instead of a bank conflict), then one range over two, then ranges below 0x80
	 * (one shared half-select in the kernel) */
	for (int pass = 0; pass < 8; pass++) {
		const bool bijective = pass < 4, single = (pass & 3) < 2, want_low = (pass & 1) == 0;
		for (size_t i = 1; i < cand.size(); i++) {
			for (size_t j = 0; j < (single ? 1 : cand.size()); j++) {
				if (j == i || (!single && j == 0)) continue;
				const bool low = cand[i].hi < 0x80 && (j == 0 || cand[j].hi < 0x80);
				if (want_low != low) continue;
				int cells = 0;
				if (!ok(cand[i], cand[j], &cells)) continue;
				if (bijective && cells != nclasses) continue;
				rlo[0] = (uint8_t) cand[i].lo; rhi[0] = (uint8_t) cand[i].hi;
				rlo[1] = (uint8_t) cand[j].lo; rhi[1] = (uint8_t) cand[j].hi;
				return low ? 1u : 2u;
			}
		}
	}
	return 0;
}

/* Byte classes: symbols whose columns are identical in every row.  One pass hashes every column
 * (O(256 S)); equal hashes are confirmed against the class representative.  cls[c] = class of byte
 * c (numbered by first occurrence), rep[k] = smallest byte of class k.  Returns the class count. */
uint32_t
byte_classes(const uint32_t *t32, uint32_t S, uint8_t (&cls)[256], uint8_t (&rep)[256])
{
	uint32_t C = 0;
	std::vector<uint64_t> colhash(256, 0x9e3779b97f4a7c15ull);
	for (uint32_t st = 0; st < S; st++) {
		const uint32_t *row = t32 + (size_t) st * 256;
		for (int c = 0; c < 256; c++) {
			uint64_t h = colhash[c] ^ row[c];
			h *= 0xff51afd7ed558ccdull;
			colhash[c] = h ^ (h >> 29);
		}
	}
	for (int c = 0; c < 256; c++) {
		int found = -1;
		for (uint32_t k = 0; k < C && found < 0; k++) {
			if (colhash[rep[k]] != colhash[c]) continue;
			bool same = true;
			for (uint32_t st = 0; st < S && same; st++) {
				same = t32[(size_t) st * 256 + c] == t32[(size_t) st * 256 + rep[k]];
			}
			if (same) found = (int) k;
		}
		if (found < 0) { rep[C] = (uint8_t) c; found = (int) C; C++; }
		cls[c] = (uint8_t) found;
	}
	return C;
}

} // namespace

/* device >= 0: validate, lay out, upload.  device < 0: validate and lay out only (the plan). */
static int
compile_impl(const struct fsm_b200_desc *desc, int device, fsm_b200_dfa **out)
{
	const bool upload = device >= 0;
	if (desc == nullptr || out == nullptr || (desc->reserved & ~FSM_B200_DESC_EAGER) != 0) {
		set_error("dfa_compile: bad argument");
		errno = EINVAL;
		return -1;
	}
	*out = nullptr;
	if (!desc_is_dfa(desc)) {
		set_error("dfa_compile: not a DFA (no start state, epsilon edge or ambiguous symbol)");
		errno = EINVAL;
		return -1;
	}

	const uint32_t S = desc->nstates;
	uint32_t *t32 = static_cast<uint32_t *>(malloc(sizeof(uint32_t) * (size_t) S * 256));
	if (t32 == nullptr) {
		errno = ENOMEM;
		return -1;
	}
	bool complete = true;
	for (uint32_t s = 0; s < S; s++) {
		uint32_t *row = t32 + (size_t) s * 256;
		for (int c = 0; c < 256; c++) row[c] = NO_EDGE;
		/* groups are disjoint for a DFA, so "first group wins" needs no ordering care */
		for (uint64_t g = desc->group_off[s]; g < desc->group_off[s + 1]; g++) {
			const uint64_t *sym = desc->group_symbols + 4 * g;
			const uint32_t to = desc->group_to[g];
			if (to >= S) {
				free(t32);
				set_error("dfa_compile: edge to state %u out of range", to);
				errno = EINVAL;
				return -1;
			}
			for (int w = 0; w < 4; w++) {
				uint64_t m = sym[w];
				while (m) {
					const int b = __builtin_ctzll(m);
					m &= m - 1;
					row[64 * w + b] = to;
				}
			}
		}
		for (int c = 0; c < 256 && complete; c++) {
			if (row[c] == NO_EDGE) complete = false;
		}
	}

	fsm_b200_dfa *dfa = new (std::nothrow) fsm_b200_dfa();
	if (dfa == nullptr) {
		free(t32);
		errno = ENOMEM;
		return -1;
	}
	memset(dfa, 0, sizeof *dfa);
	dfa->device = device;
	dfa->nstates = S;
	dfa->start = desc->start;
	dfa->complete = complete ? 1u : 0u;
	dfa->ntable = S + (complete ? 0u : 1u);
	dfa->dead = complete ? NO_EDGE : S;
	dfa->h_table32 = t32;
	dfa->entry_bytes = dfa->ntable <= 256 ? 1u : (dfa->ntable <= 65536 ? 2u : 4u);

	/* Layout choice:
	 *   dense + shared memory   rows of 256 entries, padded by 4 B      (small DFAs: K1, K1b)
	 *   classed + shared memory rows of C entries, C = byte classes      (mid-size DFAs)
	 *   classed + global (L2)   same rows, read with ld.global.nc        (large DFAs)
	 *   dense + global          when there are too many classes to gain */
	uint32_t row_pad = SMEM_ROW_PAD;
	if (const char *e = getenv("FSM_B200_ROW_PAD")) {      /* tuning knob, see DESIGN.md */
		const int v = atoi(e);
		if (v >= 0 && v <= 64 && (v % 4) == 0) row_pad = (uint32_t) v;
	}
	const uint32_t eb = dfa->entry_bytes;
	const uint64_t end_pad = (dfa->ntable + 15u) & ~15ull;
	const uint64_t dense_pitch = 256ull * eb + row_pad;
	uint32_t C = 0;
	for (int c = 0; c < 256; c++) dfa->class_of[c] = (uint8_t) c;
	if (dense_pitch * dfa->ntable + end_pad <= SMEM_TABLE_MAX && getenv("FSM_B200_FORCE_CLASSED") == nullptr) {
		dfa->smem_resident = 1;
		dfa->pitch = (uint32_t) dense_pitch;
	} else {
		uint8_t rep[256];
		C = byte_classes(t32, S, dfa->class_of, rep);
		uint64_t cpitch = ((uint64_t) C * eb + 3u) & ~3ull;
		if (((cpitch >> 2) & 1u) == 0) cpitch += 4;          /* odd word pitch spreads rows over banks */
		if (cpitch * dfa->ntable + end_pad + 256 <= SMEM_CLASS_TABLE_MAX) {
			dfa->smem_resident = 1;
			dfa->nclasses = C;
			dfa->pitch = (uint32_t) cpitch;
		} else if (C <= CLASS_GLOBAL_MAX) {
			dfa->smem_resident = 0;
			dfa->nclasses = C;
			dfa->pitch = (uint32_t) (((uint64_t) C * eb + 3u) & ~3ull);
		} else {
			dfa->smem_resident = 0;
			dfa->pitch = 256u * eb;
		}
	}
	const uint32_t ncols = dfa->nclasses ? dfa->nclasses : 256u;
	dfa->table_bytes = (uint64_t) dfa->pitch * dfa->ntable;
	const uint64_t table_pad = (dfa->table_bytes + 15u) & ~15ull;
	dfa->is_end_off = (uint32_t) table_pad;
	dfa->cls_off = (uint32_t) (table_pad + end_pad);
	dfa->blob_bytes = table_pad + end_pad + (dfa->nclasses ? 256u : 0u);
	if (dfa->blob_bytes >= (1ull << 32)) {      /* kernel-side offsets are 32-bit */
		fsm_b200_dfa_free(dfa);
		set_error("dfa_compile: transition table of %llu bytes is not supported (>= 4 GiB)", (unsigned long long) dfa->blob_bytes);
		errno = ENOTSUP;
		return -1;
	}

	std::vector<uint8_t> blob;
	try {
		blob.assign(dfa->blob_bytes, 0);
	} catch (...) {
		fsm_b200_dfa_free(dfa);
		errno = ENOMEM;
		return -1;
	}
	dfa->h_is_end = static_cast<uint8_t *>(calloc(dfa->ntable + 1, 1));
	if (dfa->h_is_end == nullptr) {
		fsm_b200_dfa_free(dfa);
		errno = ENOMEM;
		return -1;
	}
	int col_byte[256];                       /* a representative byte of every column */
	for (int c = 255; c >= 0; c--) col_byte[dfa->nclasses ? dfa->class_of[c] : c] = c;
	for (uint32_t s = 0; s < dfa->ntable; s++) {
		uint8_t *row = blob.data() + (size_t) s * dfa->pitch;
		for (uint32_t k = 0; k < ncols; k++) {
			uint32_t v = (s < S) ? t32[(size_t) s * 256 + col_byte[k]] : dfa->dead;   /* dead row absorbs */
			if (v == NO_EDGE) v = dfa->dead;
			switch (dfa->entry_bytes) {
			case 1: row[k] = (uint8_t) v; break;
			case 2: reinterpret_cast<uint16_t *>(row)[k] = (uint16_t) v; break;
			default: reinterpret_cast<uint32_t *>(row)[k] = v; break;
			}
		}
		const uint8_t e = (s < S && desc->is_end[s]) ? 1 : 0;
		dfa->h_is_end[s] = e;
		blob[dfa->is_end_off + s] = e;
	}
	if (dfa->nclasses) memcpy(blob.data() + dfa->cls_off, dfa->class_of, 256);

	if (upload) {
		FSMB_CUDA(cudaSetDevice(device), { fsm_b200_dfa_free(dfa); return -1; });
		FSMB_CUDA(cudaMalloc(&dfa->d_blob, dfa->blob_bytes), { fsm_b200_dfa_free(dfa); return -1; });
		FSMB_CUDA(cudaMemcpy(dfa->d_blob, blob.data(), dfa->blob_bytes, cudaMemcpyHostToDevice),
		    { fsm_b200_dfa_free(dfa); return -1; });
	}

	{   /* absorbing states (all 256 edges are self-loops; also the dead row): a stream chunk entered
	     * in such a state leaves in it, so K1b needs no scan job for it */
		std::vector<uint8_t> ab(dfa->ntable, 0);
		for (uint32_t st = 0; st < dfa->ntable; st++) {
			bool self = true;
			for (int c = 0; c < 256 && self; c++) {
				const uint32_t v = st < S ? t32[(size_t) st * 256 + c] : dfa->dead;
				self = (v == st);
			}
			ab[st] = self ? 1 : 0;
			if (self && st < S) dfa->has_absorbing = 1;
		}
		if (upload) {
			FSMB_CUDA(cudaMalloc(&dfa->d_absorb, dfa->ntable), { fsm_b200_dfa_free(dfa); return -1; });
			FSMB_CUDA(cudaMemcpy(dfa->d_absorb, ab.data(), dfa->ntable, cudaMemcpyHostToDevice), { fsm_b200_dfa_free(dfa); return -1; });
		}
	}

	/* ---- k-stride form: with C byte classes and C^K <= 256, index rows by the class tuple of
	 * K consecutive bytes: ONE dependent table lookup per K input bytes.  The K class lookups
	 * (256-byte LUTs, pre-multiplied by C^j) are independent of the state, conflict-free for
	 * ASCII and off the dependent chain; the shared-memory wavefronts per byte drop from
	 * 1 + conflicts to (K + 1 + conflicts) / K. */
	if (dfa->ntable <= 256 && getenv("FSM_B200_NO_KSTRIDE") == nullptr) {
		uint8_t kcls[256], rep[256];
		bool rep_ok[256];
		uint32_t KC = byte_classes(t32, S, kcls, rep);
		for (int c = 0; c < 256; c++) rep_ok[c] = (uint32_t) c < KC;
		/* ALU classification (k1_kstride_kernel, RNG != 0): when two byte ranges R0, R1 exist such
		 * that the byte class is a function of the 2-bit cell code [b in R0] + 2 [b in R1], the
		 * kernel derives the code of 4 bytes at once with integer SIMD-in-register arithmetic
		 * instead of 4 shared-memory LUT reads; the tables are then indexed by cell codes (4 per
		 * byte, some possibly unused) instead of by byte classes. */
		uint32_t krange = 0;
		uint8_t rlo[2] = { 1, 1 }, rhi[2] = { 0, 0 };          /* lo > hi: empty range */
		if (KC <= 4 && KC >= 2 && getenv("FSM_B200_NO_KRANGE") == nullptr) {
			krange = find_cell_ranges(kcls, rlo, rhi);
			if (krange != 0) {
				uint8_t cell[256];
				bool have[4] = { false, false, false, false };
				for (int c = 0; c < 256; c++) {
					cell[c] = (uint8_t) (((c >= rlo[0] && c <= rhi[0]) ? 1 : 0) | ((c >= rlo[1] && c <= rhi[1]) ? 2 : 0));
				}
				for (int c = 255; c >= 0; c--) { rep[cell[c]] = (uint8_t) c; have[cell[c]] = true; }
				for (int k = 0; k < 4; k++) rep_ok[k] = have[k];
				memcpy(kcls, cell, 256);
				KC = 4;
			}
		}
		uint32_t K = 0;
		/* K = 4 always pays.  K = 2 trades one table read for two LUT reads per 2 bytes: a win
		 * only where the table reads conflict (many live states); small DFAs (the 8-state
		 * UTF-8 validator: 2.20 vs 1.96 TB/s as a stream) stay on the one-byte kernel.
		 * (KC <= 4 <=> KC^4 <= 256, KC <= 16 <=> KC^2 <= 256: no 32-bit overflow for KC = 256.) */
		if (KC <= 4) K = 4; else if (KC <= 16 && dfa->ntable > 32) K = 2;
		if (const char *e = getenv("FSM_B200_KSTRIDE")) { const int v = atoi(e); if ((v == 2 && KC <= 16) || v == 0) K = (uint32_t) v; }
		if (K != 4) krange = 0;
		if (K != 0) {
			const uint32_t T = dfa->ntable;
			uint32_t W = 1;
			for (uint32_t j = 0; j < K; j++) W *= KC;
			if (W > 256) {          /* cannot happen: K was chosen so that KC^K <= 256 */
				fsm_b200_dfa_free(dfa);
				set_error("dfa_compile: internal error: k-stride tuple space %u", W);
				errno = EINVAL;
				return -1;
			}
			auto odd_pitch = [](uint32_t nbytes) { uint32_t p = (nbytes + 3u) & ~3u; if (((p >> 2) & 1u) == 0) p += 4; return p; };
			const uint32_t kpitch = odd_pitch(W), k1pitch = odd_pitch(KC);
			/* blob: [K class LUTs of 256 B][stepK rows][step1 rows][is_end]; the LUTs and stepK sit
			 * at compile-time offsets so every lookup is LDS [reg + uniform base + immediate] */
			const uint32_t klut_off = 0, ktab_off = 256u * K;
			const uint32_t k1_off = (ktab_off + T * kpitch + 15u) & ~15u;
			const uint32_t kend_off = (k1_off + T * k1pitch + 15u) & ~15u;
			const uint32_t kbytes = (kend_off + T + 15u) & ~15u;
			std::vector<uint8_t> kb(kbytes, 0);
			auto step1 = [&](uint32_t st, uint32_t cls) -> uint32_t {
				if (st >= S || !rep_ok[cls]) return dfa->dead == NO_EDGE ? st : dfa->dead;   /* unused cell code: never looked up */
				const uint32_t v = t32[(size_t) st * 256 + rep[cls]];
				return v == NO_EDGE ? dfa->dead : v;
			};
			for (uint32_t st = 0; st < T; st++) {
				for (uint32_t c = 0; c < KC; c++) kb[k1_off + st * k1pitch + c] = (uint8_t) step1(st, c);
				for (uint32_t idx = 0; idx < W; idx++) {
					uint32_t cur = st, rem = idx;
					for (uint32_t j = 0; j < K; j++) { cur = step1(cur, rem % KC); rem /= KC; }   /* byte j has weight C^j */
					kb[ktab_off + st * kpitch + idx] = (uint8_t) cur;
				}
				kb[kend_off + st] = dfa->h_is_end[st];
			}
			uint32_t wgt = 1;
			for (uint32_t j = 0; j < K; j++) {
				for (int c = 0; c < 256; c++) kb[klut_off + 256u * j + c] = (uint8_t) (kcls[c] * wgt);
				wgt *= KC;
			}
			if (upload) {
				FSMB_CUDA(cudaMalloc(&dfa->d_kblob, kbytes), { fsm_b200_dfa_free(dfa); return -1; });
				FSMB_CUDA(cudaMemcpy(dfa->d_kblob, kb.data(), kbytes, cudaMemcpyHostToDevice), { fsm_b200_dfa_free(dfa); return -1; });
			}
			dfa->kstride = K; dfa->kclasses = KC; dfa->kpitch = kpitch; dfa->k1pitch = k1pitch;
			dfa->k1_off = k1_off; dfa->kend_off = kend_off; dfa->klut_off = klut_off; dfa->kblob_bytes = kbytes;
			dfa->krange = krange;
			if (krange != 0) {
				/* the k-range kernel's own blob: stepK stored [tuple][state] so that the lookup address is
				 * 2 * (128 * tuple) + state -- one LEA from the dp4a result */
				const uint32_t r_k1_off = 256u + 256u * 256u;
				const uint32_t r_kend_off = (r_k1_off + T * k1pitch + 15u) & ~15u;
				const uint32_t rbytes = (r_kend_off + T + 15u) & ~15u;
				std::vector<uint8_t> rb(rbytes, 0);
				memcpy(rb.data(), kb.data() + klut_off, 256);                       /* L0: byte -> cell code */
				for (uint32_t st = 0; st < T; st++) {
					for (uint32_t idx = 0; idx < W; idx++) rb[256u + idx * 256u + st] = kb[ktab_off + st * kpitch + idx];
					memcpy(rb.data() + r_k1_off + st * k1pitch, kb.data() + k1_off + st * k1pitch, k1pitch);
					rb[r_kend_off + st] = dfa->h_is_end[st];
				}
				if (upload) {
					FSMB_CUDA(cudaMalloc(&dfa->d_rblob, rbytes), { fsm_b200_dfa_free(dfa); return -1; });
					FSMB_CUDA(cudaMemcpy(dfa->d_rblob, rb.data(), rbytes, cudaMemcpyHostToDevice), { fsm_b200_dfa_free(dfa); return -1; });
				}
				dfa->rblob_bytes = rbytes; dfa->r_k1_off = r_k1_off; dfa->r_kend_off = r_kend_off;
			}
			for (int k = 0; k < 2; k++) {
				/* per byte lane, on l = b & 0x7F: bit 7 of l + add_lo is [l >= lo7], of l + add_hi is
				 * [l > hi7]; hxor selects the half the range lives in (all ones: bytes < 0x80) */
				const bool empty = rlo[k] > rhi[k];
				const uint32_t lo7 = rlo[k] & 0x7Fu, hi7 = rhi[k] & 0x7Fu;
				dfa->kr_add_lo[k] = empty ? 0u : 0x01010101u * (0x80u - lo7);
				dfa->kr_add_hi[k] = empty ? 0u : 0x01010101u * (0x7Fu - hi7);
				dfa->kr_hxor[k] = (!empty && rlo[k] >= 0x80u) ? 0u : 0xFFFFFFFFu;
				dfa->kr_lo[k] = rlo[k]; dfa->kr_hi[k] = rhi[k];
			}
		}
	}
	/* ---- eager outputs (fsm_b200_desc_ext): dense id numbering + one bit mask per table row ---- */
	{
		const uint64_t *xoff; const uint32_t *xids;
		if (eagerhost::eh_get(desc, &xoff, &xids)) {
			std::vector<uint32_t> idl;
			std::vector<uint64_t> masks;
			eagerhost::eh_id_list(S, xoff, xids, idl);
			if (idl.size() > FSM_B200_EAGER_MAX_IDS) {
				fsm_b200_dfa_free(dfa);
				set_error("dfa_compile: %zu distinct eager-output ids (at most %u are supported)", idl.size(), FSM_B200_EAGER_MAX_IDS);
				errno = ENOTSUP;
				return -1;
			}
			dfa->eager_nbits = (uint32_t) idl.size();
			dfa->eager_words = (dfa->eager_nbits + 63u) / 64u;
			dfa->h_eager_ids = static_cast<uint32_t *>(malloc(sizeof(uint32_t) * idl.size()));
			if (dfa->h_eager_ids == nullptr) { fsm_b200_dfa_free(dfa); errno = ENOMEM; return -1; }
			memcpy(dfa->h_eager_ids, idl.data(), sizeof(uint32_t) * idl.size());
			eagerhost::eh_build_masks(S, dfa->ntable, xoff, xids, idl, dfa->eager_words, masks);
			if (upload) {
				FSMB_CUDA(cudaMalloc(&dfa->d_eager_masks, masks.size() * sizeof(uint64_t)), { fsm_b200_dfa_free(dfa); return -1; });
				FSMB_CUDA(cudaMemcpy(dfa->d_eager_masks, masks.data(), masks.size() * sizeof(uint64_t), cudaMemcpyHostToDevice),
				    { fsm_b200_dfa_free(dfa); return -1; });
			}
		}
	}
	/* ---- lines kernel blob (k1_lines.cu) ---- */
	if (dfa->ntable <= 32768 && getenv("FSM_B200_NO_LINES") == nullptr) {
		uint8_t lcls[256], lrep[256];
		const uint32_t LC = byte_classes(t32, S, lcls, lrep);
		const uint32_t T = dfa->ntable;
		const uint32_t leb = 2u;          /* 16-bit entries: the kernel rewrites them into row handles (k1_lines.cu) */
		uint32_t lpitch = ((LC + 1) * leb + 3u) & ~3u;
		if (((lpitch >> 2) & 1u) == 0) lpitch += 4;                /* odd word pitch spreads rows over banks */
		const uint32_t ltab_off = 512;
		const uint32_t lend_off = (ltab_off + T * lpitch + 15u) & ~15u;
		const uint64_t lbytes = ((uint64_t) lend_off + T + 15u) & ~15ull;
		if ((LC + 1) * leb <= 256 && lbytes <= SMEM_LINES_MAX) {       /* the LUT holds class * entry size in a byte */
			/* new numbering: plain states, then states with eager outputs, then the dead row */
			std::vector<uint32_t> perm(T), inv(T);
			std::vector<uint64_t> hmask;       /* host copy of the row masks, for "has outputs" */
			const uint64_t *xoff; const uint32_t *xids;
			const bool has_eager = eagerhost::eh_get(desc, &xoff, &xids);
			uint32_t nplain = 0;
			for (uint32_t st = 0; st < S; st++) if (!has_eager || xoff[st + 1] == xoff[st]) { perm[st] = nplain; inv[nplain++] = st; }
			uint32_t nn = nplain;
			for (uint32_t st = 0; st < S; st++) if (has_eager && xoff[st + 1] != xoff[st]) { perm[st] = nn; inv[nn++] = st; }
			if (!complete) { perm[S] = S; inv[S] = S; }
			std::vector<uint8_t> lb(lbytes, 0);
			for (int c = 0; c < 256; c++) { lb[c] = (uint8_t) (lcls[c] * leb); lb[256 + c] = (uint8_t) (LC * leb); }
			for (uint32_t ns = 0; ns < T; ns++) {
				const uint32_t os = inv[ns];
				uint8_t *row = lb.data() + ltab_off + (size_t) ns * lpitch;
				for (uint32_t k = 0; k <= LC; k++) {
					uint32_t v;
					if (k == LC) v = ns;                                        /* NOP column */
					else if (os >= S) v = perm[dfa->dead];                      /* dead row absorbs */
					else {
						const uint32_t t = t32[(size_t) os * 256 + lrep[k]];
						v = perm[t == NO_EDGE ? dfa->dead : t];
					}
					if (leb == 1) row[k] = (uint8_t) v; else reinterpret_cast<uint16_t *>(row)[k] = (uint16_t) v;
				}
				lb[lend_off + ns] = dfa->h_is_end[os];
			}
			dfa->lblob_bytes = (uint32_t) lbytes; dfa->l_pitch = lpitch; dfa->l_entry_bytes = leb;
			dfa->l_end_off = lend_off; dfa->l_ncols = LC + 1; dfa->l_tab_off = ltab_off;
			dfa->l_first_event = (has_eager || !complete) ? nplain : NO_EDGE;
			dfa->l_dead = complete ? NO_EDGE : S;
			dfa->l_start = perm[dfa->start];
			if (has_eager && dfa->eager_words <= 4) {
				std::vector<uint32_t> idl;
				eagerhost::eh_id_list(S, xoff, xids, idl);
				eagerhost::eh_build_masks(S, dfa->ntable, xoff, xids, idl, dfa->eager_words, hmask);
				for (uint32_t w = 0; w < dfa->eager_words; w++) dfa->l_start_mask[w] = hmask[(size_t) dfa->start * dfa->eager_words + w];
			}
			if (upload) {
				FSMB_CUDA(cudaMalloc(&dfa->d_lblob, lbytes), { fsm_b200_dfa_free(dfa); return -1; });
				FSMB_CUDA(cudaMemcpy(dfa->d_lblob, lb.data(), lbytes, cudaMemcpyHostToDevice), { fsm_b200_dfa_free(dfa); return -1; });
				FSMB_CUDA(cudaMalloc(&dfa->d_lperm_inv, T * sizeof(uint32_t)), { fsm_b200_dfa_free(dfa); return -1; });
				FSMB_CUDA(cudaMemcpy(dfa->d_lperm_inv, inv.data(), T * sizeof(uint32_t), cudaMemcpyHostToDevice), { fsm_b200_dfa_free(dfa); return -1; });
				FSMB_CUDA(cudaMalloc(&dfa->d_lperm, T * sizeof(uint32_t)), { fsm_b200_dfa_free(dfa); return -1; });
				FSMB_CUDA(cudaMemcpy(dfa->d_lperm, perm.data(), T * sizeof(uint32_t), cudaMemcpyHostToDevice), { fsm_b200_dfa_free(dfa); return -1; });
				if (has_eager && dfa->eager_words <= 4 && !hmask.empty()) {
					const uint32_t W = dfa->eager_words, nev = T - nplain;
					std::vector<uint64_t> ev((size_t) nev * W + 1, 0);
					for (uint32_t ns = nplain; ns < T; ns++) {
						for (uint32_t w = 0; w < W; w++) ev[(size_t) (ns - nplain) * W + w] = hmask[(size_t) inv[ns] * W + w];
					}
					FSMB_CUDA(cudaMalloc(&dfa->d_lev_masks, ev.size() * sizeof(uint64_t)), { fsm_b200_dfa_free(dfa); return -1; });
					FSMB_CUDA(cudaMemcpy(dfa->d_lev_masks, ev.data(), ev.size() * sizeof(uint64_t), cudaMemcpyHostToDevice), { fsm_b200_dfa_free(dfa); return -1; });
				}
				if (dfa->has_absorbing) {
					std::vector<uint8_t> lab(T, 0);
					for (uint32_t ns = 0; ns < T; ns++) {
						const uint32_t os = inv[ns];
						bool self = os < S;
						for (int c = 0; c < 256 && self; c++) self = t32[(size_t) os * 256 + c] == os;
						lab[ns] = self ? 1 : 0;
					}
					FSMB_CUDA(cudaMalloc(&dfa->d_labsorb, T), { fsm_b200_dfa_free(dfa); return -1; });
					FSMB_CUDA(cudaMemcpy(dfa->d_labsorb, lab.data(), T, cudaMemcpyHostToDevice), { fsm_b200_dfa_free(dfa); return -1; });
				}
			} else {
				dfa->d_lblob = nullptr;
			}
			dfa->l_planned = 1;
		}
	}
	*out = dfa;
	return 0;
}

extern "C" int
fsm_b200_dfa_eager_info(const fsm_b200_dfa *dfa, uint32_t *nbits, const uint32_t **id_of_bit)
{
	if (dfa == nullptr || nbits == nullptr || id_of_bit == nullptr) {
		errno = EINVAL;
		return -1;
	}
	*nbits = dfa->eager_nbits;
	*id_of_bit = dfa->h_eager_ids;
	return 0;
}

extern "C" int
fsm_b200_dfa_compile(const struct fsm_b200_desc *desc, int device, fsm_b200_dfa **out)
{
	if (device < 0) {
		set_error("dfa_compile: bad device %d", device);
		errno = EINVAL;
		return -1;
	}
	return compile_impl(desc, device, out);
}

namespace fsmb200 { void scratch_free(fsm_b200_dfa *dfa); void stream_scratch_free(fsm_b200_dfa *dfa); void eager_scratch_free(fsm_b200_dfa *dfa); }

/* The layout a compile WOULD choose, without touching any device: host logic only. */
extern "C" int
fsm_b200_dfa_plan(const struct fsm_b200_desc *desc, struct fsm_b200_dfa_info *info)
{
	fsm_b200_dfa *dfa = nullptr;
	if (info == nullptr) {
		errno = EINVAL;
		return -1;
	}
	if (compile_impl(desc, -1, &dfa) != 0) return -1;
	const int rc = fsm_b200_dfa_info(dfa, info);
	info->device = 0xFFFFFFFFu;
	fsm_b200_dfa_free(dfa);            /* nothing was allocated on a device: frees host copies only */
	return rc;
}

extern "C" void
fsm_b200_dfa_free(fsm_b200_dfa *dfa)
{
	if (dfa == nullptr) return;
	fsmb200::scratch_free(dfa);
	fsmb200::stream_scratch_free(dfa);
	fsmb200::eager_scratch_free(dfa);
	if (dfa->d_blob != nullptr) {
		cudaSetDevice(dfa->device);
		cudaFree(dfa->d_blob);
	}
	if (dfa->d_kblob != nullptr) cudaFree(dfa->d_kblob);
	if (dfa->d_rblob != nullptr) cudaFree(dfa->d_rblob);
	if (dfa->d_absorb != nullptr) cudaFree(dfa->d_absorb);
	if (dfa->d_eager_masks != nullptr) cudaFree(dfa->d_eager_masks);
	if (dfa->d_lblob != nullptr) cudaFree(dfa->d_lblob);
	if (dfa->d_lperm_inv != nullptr) cudaFree(dfa->d_lperm_inv);
	if (dfa->d_lperm != nullptr) cudaFree(dfa->d_lperm);
	if (dfa->d_labsorb != nullptr) cudaFree(dfa->d_labsorb);
	if (dfa->d_lev_masks != nullptr) cudaFree(dfa->d_lev_masks);
	free(dfa->h_eager_ids);
	free(dfa->h_table32);
	free(dfa->h_is_end);
	delete dfa;
}

extern "C" int
fsm_b200_dfa_info(const fsm_b200_dfa *dfa, struct fsm_b200_dfa_info *info)
{
	if (dfa == nullptr || info == nullptr) {
		errno = EINVAL;
		return -1;
	}
	info->nstates = dfa->nstates;
	info->ntable_states = dfa->ntable;
	info->start = dfa->start;
	info->entry_bytes = dfa->entry_bytes;
	info->row_pitch_bytes = dfa->pitch;
	info->complete = dfa->complete;
	info->smem_resident = dfa->smem_resident;
	info->device = (uint32_t) dfa->device;
	info->table_bytes = dfa->table_bytes;
	info->nclasses = dfa->nclasses;
	info->kstride = dfa->kstride;
	info->krange = dfa->krange;
	for (int k = 0; k < 2; k++) { info->krange_lo[k] = dfa->kr_lo[k]; info->krange_hi[k] = dfa->kr_hi[k]; }
	info->kclasses = dfa->kclasses;
	info->lines_smem = dfa->l_planned;
	info->lines_blob_bytes = dfa->l_planned ? dfa->lblob_bytes : 0;
	info->lines_cols = dfa->l_planned ? dfa->l_ncols : 0;
	info->eager_ids = dfa->eager_nbits;
	return 0;
}

/* Read the table back FROM THE DEVICE (so tests see what the kernels see). */
extern "C" int
fsm_b200_dfa_table(const fsm_b200_dfa *dfa, uint32_t *out)
{
	if (dfa == nullptr || out == nullptr) {
		errno = EINVAL;
		return -1;
	}
	std::vector<uint8_t> blob(dfa->blob_bytes);
	FSMB_CUDA(cudaSetDevice(dfa->device), return -1);
	FSMB_CUDA(cudaMemcpy(blob.data(), dfa->d_blob, dfa->blob_bytes, cudaMemcpyDeviceToHost), return -1);
	for (uint32_t s = 0; s < dfa->nstates; s++) {
		const uint8_t *row = blob.data() + (size_t) s * dfa->pitch;
		const uint8_t *cls = dfa->nclasses ? blob.data() + dfa->cls_off : nullptr;
		for (int c = 0; c < 256; c++) {
			const uint32_t k = cls ? cls[c] : (uint32_t) c;
			uint32_t v;
			switch (dfa->entry_bytes) {
			case 1: v = row[k]; break;
			case 2: v = reinterpret_cast<const uint16_t *>(row)[k]; break;
			default: v = reinterpret_cast<const uint32_t *>(row)[k]; break;
			}
			out[(size_t) s * 256 + c] = (v == dfa->dead) ? NO_EDGE : v;
		}
	}
	return 0;
}
