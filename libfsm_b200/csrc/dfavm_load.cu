/*
 * dfavm_load.cu -- loader for the reference's DFAVM bytecode: a DFA saved by
 * fsm_dfavm_save / `fsm -l vmops`-style tooling as a "DFAVM$" image (src/libfsm/vm.c:39-71, variable
 * encoding 0.1: vm/v1.c:20-49 save, :84-220 encode, :321-432 interpret) becomes a flat description the
 * engine compiles and scans like any other DFA -- without the `struct fsm` it was generated from.
 *
 * The program is a list of states, each starting with a FETCH (bit 0: accept if the input ends here)
 * followed by conditional BRANCH / STOP instructions on the fetched byte.  State i of the description is
 * the i-th FETCH (in address order; the start state is where the program begins); for every byte value the instructions after the FETCH are run
 * symbolically until the next FETCH (= the destination), STOP-fail (no edge) or STOP-success (an extra
 * absorbing accepting state: the VM stops reading there and reports a match).  Host code only.
 * The VM answers yes / no only (fsm_vm_match_buffer, vm.c:218-229): there are no end ids in the image.
 */
#include <array>
#include <cstring>
#include <map>
#include <new>
#include <vector>

#include "common.h"

using namespace fsmb200;

namespace {

enum { OP_STOP = 0, OP_FETCH = 1, OP_BRANCH = 2 };
enum { CMP_ALWAYS = 0, CMP_LT = 1, CMP_LE = 2, CMP_GE = 3, CMP_GT = 4, CMP_EQ = 5, CMP_NE = 6 };

struct Decoded { uint32_t op, cmp, arg, rest, next; int32_t rel; bool ok; };

/* one instruction at pc (vm/v1.c:331-425) */
Decoded
decode(const uint8_t *ops, uint32_t len, uint32_t pc)
{
	Decoded d; memset(&d, 0, sizeof d);
	if (pc >= len) return d;
	const uint8_t b = ops[pc];
	d.op = (b >> 3) & 3u; d.cmp = b >> 5; d.rest = b & 7u;
	uint32_t off = pc + 1;
	if (d.op == OP_FETCH) { d.next = off; d.ok = d.cmp == CMP_ALWAYS; return d; }
	if (d.cmp > CMP_NE || d.op > OP_BRANCH) return d;
	if (d.cmp != CMP_ALWAYS) { if (off >= len) return d; d.arg = ops[off++]; }
	if (d.op == OP_BRANCH) {
		const uint32_t dest = b & 3u, nb = 1u << dest;
		if (dest > 2 || off + nb > len) return d;
		if (dest == 0) d.rel = (int8_t) ops[off];
		else if (dest == 1) d.rel = (int16_t) (ops[off] | (ops[off + 1] << 8));
		else d.rel = (int32_t) ((uint32_t) ops[off] | ((uint32_t) ops[off + 1] << 8) | ((uint32_t) ops[off + 2] << 16) | ((uint32_t) ops[off + 3] << 24));
		off += nb;
	}
	d.next = off; d.ok = true;
	return d;
}

bool
holds(uint32_t cmp, int ch, int arg)
{
	switch (cmp) {
	case CMP_ALWAYS: return true;
	case CMP_LT: return ch < arg;
	case CMP_LE: return ch <= arg;
	case CMP_GE: return ch >= arg;
	case CMP_GT: return ch > arg;
	case CMP_EQ: return ch == arg;
	case CMP_NE: return ch != arg;
	default: return false;
	}
}

} // namespace

extern "C" int
fsm_b200_dfavm_load(const uint8_t *image, size_t nbytes, struct fsm_b200_owned_desc *out)
{
	if (image == nullptr || out == nullptr) { errno = EINVAL; return -1; }
	memset(out, 0, sizeof *out);
	if (nbytes < 12 || memcmp(image, "DFAVM$", 6) != 0) {
		set_error("dfavm_load: not a DFAVM image");
		errno = EINVAL;
		return -1;
	}
	if (image[6] != 0x00 || image[7] != 0x01) {
		set_error("dfavm_load: unsupported encoding %u.%u (only the variable encoding 0.1 is)", image[6], image[7]);
		errno = ENOTSUP;
		return -1;
	}
	const uint32_t len = (uint32_t) image[8] | ((uint32_t) image[9] << 8) | ((uint32_t) image[10] << 16) | ((uint32_t) image[11] << 24);
	if ((size_t) len + 12 > nbytes || len == 0) {
		set_error("dfavm_load: truncated image (%u instruction bytes announced, %zu present)", len, nbytes - 12);
		errno = EINVAL;
		return -1;
	}
	const uint8_t *ops = image + 12;

	/* pass 1: instruction boundaries; every FETCH is a state */
	std::map<uint32_t, uint32_t> state_of;          /* address of a FETCH -> state number */
	std::vector<uint32_t> fetch_at;
	for (uint32_t pc = 0; pc < len; ) {
		const Decoded d = decode(ops, len, pc);
		if (!d.ok) {
			set_error("dfavm_load: bad instruction at offset %u", pc);
			errno = EINVAL;
			return -1;
		}
		if (d.op == OP_FETCH) { state_of[pc] = (uint32_t) fetch_at.size(); fetch_at.push_back(pc); }
		pc = d.next;
	}
	const uint32_t nfetch = (uint32_t) fetch_at.size();
	const uint32_t ACCEPT_ALL = nfetch;             /* STOP-success: absorbing, accepting */
	bool need_sink = false;
	/* where the program starts: normally its first instruction is a FETCH; an automaton that accepts (or
	 * rejects) everything compiles to a bare STOP, which the VM runs before reading anything (ch = 0) */
	uint32_t start = NO_EDGE;
	bool reject_all = false;
	{
		uint32_t pc = 0, steps = 0;
		for (;;) {
			if (++steps > 4096 || pc >= len) { set_error("dfavm_load: control flow leaves the program at offset %u", pc); errno = EINVAL; return -1; }
			const Decoded d = decode(ops, len, pc);
			if (d.op == OP_FETCH) { start = state_of[pc]; break; }
			if (holds(d.cmp, 0, (int) d.arg)) {
				if (d.op == OP_STOP) { if (d.rest & 1u) { start = ACCEPT_ALL; need_sink = true; } else reject_all = true; break; }
				pc = (uint32_t) ((int64_t) pc + d.rel);
			} else {
				pc = d.next;
			}
		}
	}

	Owner *ow = new (std::nothrow) Owner();
	if (ow == nullptr) { errno = ENOMEM; return -1; }
	std::vector<uint32_t> dst(256);
	ow->group_off.push_back(0);
	for (uint32_t s = 0; s < nfetch; s++) {
		const Decoded f = decode(ops, len, fetch_at[s]);
		ow->is_end.push_back((uint8_t) (f.rest & 1u));
		for (int ch = 0; ch < 256; ch++) {
			uint32_t pc = f.next, steps = 0;
			uint32_t to = NO_EDGE;
			for (;;) {
				if (++steps > 4096 || pc >= len) { delete ow; set_error("dfavm_load: control flow leaves the program at offset %u", pc); errno = EINVAL; return -1; }
				const Decoded d = decode(ops, len, pc);
				if (!d.ok) { delete ow; set_error("dfavm_load: bad instruction at offset %u", pc); errno = EINVAL; return -1; }
				if (d.op == OP_FETCH) {
					to = state_of[pc];
					break;
				}
				if (holds(d.cmp, ch, (int) d.arg)) {
					if (d.op == OP_STOP) { if (d.rest & 1u) { to = ACCEPT_ALL; need_sink = true; } break; }
					pc = (uint32_t) ((int64_t) pc + d.rel);              /* relative to the branch itself (v1.c:412-416) */
				} else {
					pc = d.next;
				}
			}
			dst[ch] = to;
		}
		/* groups: one per distinct destination, ascending (the order edge_set keeps) */
		std::map<uint32_t, std::array<uint64_t, 4>> groups;
		for (int ch = 0; ch < 256; ch++) {
			if (dst[ch] == NO_EDGE) continue;
			auto &m = groups[dst[ch]];
			m[ch >> 6] |= 1ull << (ch & 63);
		}
		for (auto &g : groups) {
			ow->group_to.push_back(g.first);
			for (int w = 0; w < 4; w++) ow->group_sym.push_back(g.second[w]);
		}
		ow->group_off.push_back(ow->group_to.size());
	}
	if (reject_all) {                               /* one state, not accepting, no edges */
		start = (uint32_t) ow->is_end.size();
		ow->is_end.push_back(0);
		ow->group_off.push_back(ow->group_to.size());
	}
	if (need_sink) {
		ow->is_end.push_back(1);
		ow->group_to.push_back(ACCEPT_ALL);
		for (int w = 0; w < 4; w++) ow->group_sym.push_back(~0ull);
		ow->group_off.push_back(ow->group_to.size());
	}
	ow->endid_off.assign(ow->is_end.size() + 1, 0);
	ow->endids.push_back(0);
	out->owner = ow;
	out->desc.nstates = (uint32_t) ow->is_end.size();
	out->desc.start = start == ACCEPT_ALL ? (uint32_t) ow->is_end.size() - 1u : start; out->desc.hasstart = 1; out->desc.reserved = 0;
	out->desc.is_end = ow->is_end.data();
	out->desc.group_off = ow->group_off.data();
	out->desc.group_symbols = ow->group_sym.data();
	out->desc.group_to = ow->group_to.data();
	out->desc.eps_off = nullptr; out->desc.eps_to = nullptr;
	out->desc.endid_off = ow->endid_off.data(); out->desc.endids = ow->endids.data();
	return 0;
}
