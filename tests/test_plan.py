"""CPU: the flattener's host logic -- validation and table-layout decisions -- through
fsm_b200_dfa_plan (no device needed), on the golden automata recorded from the reference."""
import errno
import os

import pytest

import goldenio
import libfsm_b200 as L

CASES = {c["name"]: c for c in goldenio.load_exec_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_exec.npz"))}


def case(prefix):
    return next(c for n, c in CASES.items() if n.startswith(prefix))


def test_config2_dfa_layout():
    p = L.plan(case("cfg2:uniform")["fsm"])
    assert p["nstates"] == 256 and p["ntable_states"] == 256 and p["complete"] == 1   # no dead row needed
    assert p["entry_bytes"] == 1 and p["row_pitch_bytes"] == 260 and p["smem_resident"] == 1
    assert p["nclasses"] == 0 and p["kstride"] == 4                                     # 3 byte classes -> K = 4
    assert p["table_bytes"] == 256 * 260


def test_incomplete_dfa_gets_a_dead_row():
    p = L.plan(case("anchored:^abc[0-9]+x$")["fsm"])
    assert p["complete"] == 0 and p["ntable_states"] == p["nstates"] + 1
    assert p["entry_bytes"] == 1 and p["smem_resident"] == 1 and p["kstride"] == 0     # small table: one-byte kernel


def test_wide_tables_use_byte_class_rows():
    f = case("union6:")["fsm"]
    p = L.plan(f)
    assert f.nstates > 256 and p["entry_bytes"] == 2
    assert 1 <= p["nclasses"] <= 64 and p["smem_resident"] == 1                        # classed rows fit shared memory
    assert p["row_pitch_bytes"] >= 2 * p["nclasses"] and (p["row_pitch_bytes"] // 4) % 2 == 1   # odd word pitch
    assert p["kstride"] == 0


def test_kstride_choices():
    assert L.plan(case("cfg1:digits")["fsm"])["kstride"] == 4
    assert L.plan(case("endids:union6x5")["fsm"])["kstride"] == 2                        # <= 16 classes, > 32 rows
    assert L.plan(case("utf8:")["fsm"])["kstride"] == 0                                  # 8 states: stays one-byte


@pytest.mark.parametrize("name", sorted(CASES))
def test_plan_agrees_with_validation(oracle, name):
    c = CASES[name]
    if c["is_dfa"]:
        p = L.plan(c["fsm"])
        assert p["nstates"] == c["fsm"].nstates and p["start"] == c["fsm"].start
        complete = bool((oracle.flatten(c["fsm"]) != 0xFFFFFFFF).all())
        assert p["complete"] == int(complete)
    else:
        with pytest.raises(L.FsmB200Error) as ei:
            L.plan(c["fsm"])
        assert ei.value.errno == errno.EINVAL


def test_layout_knobs(monkeypatch):
    f = case("cfg2:uniform")["fsm"]
    monkeypatch.setenv("FSM_B200_ROW_PAD", "0")
    assert L.plan(f)["row_pitch_bytes"] == 256
    monkeypatch.delenv("FSM_B200_ROW_PAD")
    monkeypatch.setenv("FSM_B200_FORCE_CLASSED", "1")
    p = L.plan(f)
    assert p["nclasses"] == 3 and p["row_pitch_bytes"] == 4 and p["smem_resident"] == 1
    monkeypatch.delenv("FSM_B200_FORCE_CLASSED")
    monkeypatch.setenv("FSM_B200_KSTRIDE", "2")
    assert L.plan(f)["kstride"] == 2
    monkeypatch.setenv("FSM_B200_KSTRIDE", "0")
    assert L.plan(f)["kstride"] == 0
    monkeypatch.delenv("FSM_B200_KSTRIDE")
    monkeypatch.setenv("FSM_B200_NO_KSTRIDE", "1")
    assert L.plan(f)["kstride"] == 0


def test_plan_accepts_eager_outputs_up_to_the_id_limit():
    """A desc_ext is validated and laid out like a plain desc; more than FSM_B200_EAGER_MAX_IDS (256)
    distinct eager-output ids are refused with ENOTSUP at compile time, not at exec time."""
    import errno
    from libfsm_b200.desc import FlatFsm
    n = 300
    edges = [(s, 97, s + 1) for s in range(n - 1)]
    ok = FlatFsm.from_edges(n, 0, [n - 1], edges, eager={s: [1000 + s] for s in range(256)})
    assert L.plan(ok)["nstates"] == n
    too_many = FlatFsm.from_edges(n, 0, [n - 1], edges, eager={s: [1000 + s] for s in range(257)})
    with pytest.raises(OSError) as e:
        L.plan(too_many)
    assert e.value.errno == errno.ENOTSUP


def test_kstride_choice_does_not_overflow_with_256_byte_classes():
    """KC^4 was evaluated in 32 bits, so 256 distinct byte columns (256^4 == 2^32 -> 0)
    chose K = 4 with an empty tuple table.  18 states: A: c -> 2 + c / 16, B: c -> 2 + c % 16."""
    from libfsm_b200.desc import FlatFsm
    edges = [(0, c, 2 + c // 16) for c in range(256)] + [(1, c, 2 + c % 16) for c in range(256)]
    edges += [(s, list(range(256)), s % 2) for s in range(2, 18)]
    p = L.plan(FlatFsm.from_edges(18, 0, [17], edges))
    assert p["kstride"] == 0 and p["krange"] == 0


def _cells_refine_classes(p, fsm):
    """The plan's two ranges: every byte's column of the dense table is determined by its cell code."""
    import numpy as np
    t = fsm.dense_table()
    b = np.arange(256)
    cell = np.zeros(256, dtype=np.int64)
    for k in range(2):
        lo, hi = p["krange_lo"][k], p["krange_hi"][k]
        assert lo > hi or (lo < 0x80) == (hi < 0x80)         # a range stays inside one half of the byte space
        cell |= ((b >= lo) & (b <= hi)).astype(np.int64) << k
    for c in range(4):
        cols = t[:, cell == c]
        if cols.shape[1]:
            assert (cols == cols[:, :1]).all(), f"cell {c} mixes byte classes"
    return cell


def test_alu_classification_ranges():
    """find_cell_ranges: config 2 (printable / 'a' / rest) and config 1 (digit / '.' / rest) are each
    two ranges below 0x80; classes that need a range in the upper half give krange 2; classes that no
    two ranges separate keep the LUT kernel (krange 0)."""
    import synth
    f = case("cfg2:uniform")["fsm"]
    p = L.plan(f)
    assert p["kstride"] == 4 and p["krange"] == 1 and p["kclasses"] == 4
    cell = _cells_refine_classes(p, f)
    assert cell[0x61] not in (cell[0x20], cell[0x7E], cell[0x00]) and cell[0x7F] == cell[0x00] == cell[0xFF]
    f = case("cfg1:digits")["fsm"]
    p = L.plan(f)
    assert p["krange"] == 1
    _cells_refine_classes(p, f)
    # upper-half range
    f = synth.dfa_from_classes(synth.classes_from_ranges((0x30, 0x39), (0xC2, 0xDF)), 40, seed=1)
    p = L.plan(f)
    assert p["kstride"] == 4 and p["krange"] == 2
    _cells_refine_classes(p, f)
    # a single range is enough for two classes; the second comes back empty
    f = synth.dfa_from_classes(synth.classes_from_ranges((0x41, 0x5A)), 40, seed=2)
    p = L.plan(f)
    assert p["krange"] == 1 and (p["krange_lo"][1] > p["krange_hi"][1])
    _cells_refine_classes(p, f)
    # [a-zA-Z] vs digits vs rest: 3 classes in 6 runs -> no two ranges
    c = synth.classes_from_ranges((0x41, 0x5A), (0x30, 0x39))
    c[0x61:0x7B] = c[0x41]
    p = L.plan(synth.dfa_from_classes(c, 40, seed=3))
    assert p["kstride"] == 4 and p["krange"] == 0 and p["kclasses"] == 3
    # a range that straddles 0x80 ([0x70, 0x8F]) is two ranges for the engine: one per half
    f = synth.dfa_from_classes(synth.classes_from_ranges((0x70, 0x8F)), 40, seed=4)
    p = L.plan(f)
    assert p["kstride"] == 4 and p["krange"] == 2
    _cells_refine_classes(p, f)
