"""Load/save the golden fixture bundles under tests/golden/ (plain npz, no pickles)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from libfsm_b200.desc import FlatFsm, RESULT_DTYPE  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
FSM_KEYS = ("is_end", "group_off", "group_symbols", "group_to", "eps_off", "eps_to", "endid_off", "endids")


def pack_fsm(prefix: str, f: FlatFsm, out: dict) -> None:
    out[prefix + "hdr"] = np.array([f.nstates, f.start, int(f.hasstart)], dtype=np.int64)
    for k in FSM_KEYS:
        out[prefix + k] = getattr(f, k)
    if f.eager_off is not None:                    # eager-output sets (only golden_eager.npz has them)
        out[prefix + "eager_off"] = f.eager_off
        out[prefix + "eager_ids"] = f.eager_ids


def unpack_fsm(prefix: str, z) -> FlatFsm:
    n, start, has = (int(x) for x in z[prefix + "hdr"])
    extra = {}
    if (prefix + "eager_off") in z.files:
        extra = {"eager_off": z[prefix + "eager_off"], "eager_ids": z[prefix + "eager_ids"]}
    return FlatFsm(nstates=n, start=start, hasstart=bool(has), **{k: z[prefix + k] for k in FSM_KEYS}, **extra)


def save_exec_cases(path: str, cases: list[dict]) -> None:
    """case: name, fsm, base(u8), offsets(u64), expect(RESULT_DTYPE; end valid where ret==1),
    expect_amortised (RESULT_DTYPE, end valid everywhere) or None, tst_expect (i8: 1 '+', 0 '-', -1 n/a),
    is_dfa (bool)."""
    out, meta = {}, []
    for i, c in enumerate(cases):
        p = f"c{i}_"
        pack_fsm(p, c["fsm"], out)
        out[p + "base"] = c["base"]
        out[p + "offsets"] = c["offsets"]
        out[p + "expect"] = c["expect"].view(np.uint8)
        if c.get("expect_amortised") is not None:
            out[p + "expect_am"] = c["expect_amortised"].view(np.uint8)
        tst = c.get("tst_expect")
        if tst is None:
            tst = np.full(len(c["offsets"]) - 1, -1)
        out[p + "tst"] = np.asarray(tst, dtype=np.int8)
        meta.append({"name": c["name"], "is_dfa": bool(c.get("is_dfa", True)), "note": c.get("note", "")})
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **out)


def load_exec_cases(path: str) -> list[dict]:
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    cases = []
    for i, m in enumerate(meta):
        p = f"c{i}_"
        c = dict(m)
        c["fsm"] = unpack_fsm(p, z)
        c["base"] = z[p + "base"]
        c["offsets"] = z[p + "offsets"]
        c["expect"] = z[p + "expect"].view(RESULT_DTYPE)
        c["expect_amortised"] = z[p + "expect_am"].view(RESULT_DTYPE) if (p + "expect_am") in z.files else None
        c["tst_expect"] = z[p + "tst"]
        cases.append(c)
    return cases


def save_det_cases(path: str, cases: list[dict]) -> None:
    """case: name, nfa (FlatFsm), dfa (FlatFsm: the REFERENCE's fsm_determinise output),
    closure_off/closure_to (reference epsilon_closure CSR) or None."""
    out, meta = {}, []
    for i, c in enumerate(cases):
        p = f"d{i}_"
        pack_fsm(p + "nfa_", c["nfa"], out)
        pack_fsm(p + "dfa_", c["dfa"], out)
        if c.get("closure_off") is not None:
            out[p + "cl_off"] = c["closure_off"]; out[p + "cl_to"] = c["closure_to"]
        meta.append({"name": c["name"], "note": c.get("note", "")})
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **out)


def load_det_cases(path: str) -> list[dict]:
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    cases = []
    for i, m in enumerate(meta):
        p = f"d{i}_"
        c = dict(m)
        c["nfa"] = unpack_fsm(p + "nfa_", z)
        c["dfa"] = unpack_fsm(p + "dfa_", z)
        c["closure_off"] = z[p + "cl_off"] if (p + "cl_off") in z.files else None
        c["closure_to"] = z[p + "cl_to"] if (p + "cl_to") in z.files else None
        cases.append(c)
    return cases


def load_re_fixtures(path: str) -> list[dict]:
    """The reference's regex golden files: name, dialect, re(1) args, regex bytes, expected automaton."""
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    out = []
    for i, m in enumerate(meta):
        c = dict(m)
        c["regex"] = bytes(z[f"f{i}_regex"])
        c["fsm"] = unpack_fsm(f"f{i}_", z)
        out.append(c)
    return out


def save_eager_cases(path: str, cases: list[dict]) -> None:
    """case: name, nfa, dfa (the reference's fsm_determinise of nfa), min (its fsm_minimise of dfa, or
    None when nothing is left), inputs (list of bytes), fired (list of sorted id lists: what the
    reference's fsm_exec handed to the eager-output callback on `min`, or on `dfa` when min is None),
    rets (fsm_exec's return per input)."""
    out, meta = {}, []
    for i, c in enumerate(cases):
        p = f"e{i}_"
        pack_fsm(p + "nfa_", c["nfa"], out)
        pack_fsm(p + "dfa_", c["dfa"], out)
        if c["min"] is not None:
            pack_fsm(p + "min_", c["min"], out)
        meta.append({"name": c["name"], "has_min": c["min"] is not None,
                     "inputs": [s.hex() for s in c["inputs"]], "fired": c["fired"], "rets": c["rets"]})
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **out)


def load_eager_cases(path: str) -> list[dict]:
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    cases = []
    for i, m in enumerate(meta):
        p = f"e{i}_"
        cases.append({"name": m["name"], "nfa": unpack_fsm(p + "nfa_", z), "dfa": unpack_fsm(p + "dfa_", z),
                      "min": unpack_fsm(p + "min_", z) if m["has_min"] else None,
                      "inputs": [bytes.fromhex(h) for h in m["inputs"]], "fired": m["fired"], "rets": m["rets"]})
    return cases


def load_cfg3(path: str | None = None) -> dict:
    """golden_cfg3.npz (make_golden.py main_cfg3): {"eager": {fsm, base, offsets, expect, masks, idlist},
    "anchored": {fsm, base, offsets, expect}, "meta": {...}}."""
    z = np.load(path or os.path.join(GOLDEN_DIR, "golden_cfg3.npz"))
    return {
        "eager": {"fsm": unpack_fsm("eager_", z), "base": z["eager_base"], "offsets": z["eager_offsets"],
                  "expect": z["eager_expect"].view(RESULT_DTYPE), "masks": z["eager_masks"], "idlist": z["eager_idlist"]},
        "anchored": {"fsm": unpack_fsm("anch_", z), "base": z["anch_base"], "offsets": z["anch_offsets"],
                     "expect": z["anch_expect"].view(RESULT_DTYPE)},
        "meta": json.loads(bytes(z["meta"]).decode()),
    }


def load_cfg4(path: str | None = None) -> dict:
    """golden_cfg4.npz (make_golden.py main_cfg4): {"fsm": the utf8dfa-star validator, "meta": {...}}."""
    z = np.load(path or os.path.join(GOLDEN_DIR, "golden_cfg4.npz"))
    return {"fsm": unpack_fsm("utf8dfa_star_", z), "meta": json.loads(bytes(z["meta"]).decode())}


def load_cfg5eps(path: str | None = None) -> dict:
    """golden_cfg5eps.npz (make_golden.py main_cfg5eps): {"nfa": FlatFsm, "meta": {dfa_states, dfa_canonical_sha256, ...}}."""
    z = np.load(path or os.path.join(GOLDEN_DIR, "golden_cfg5eps.npz"))
    return {"nfa": unpack_fsm("nfa_", z), "meta": json.loads(bytes(z["meta"]).decode())}
