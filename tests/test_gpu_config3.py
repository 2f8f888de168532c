"""GPU: BASELINE config 3 shape -- an rx(1)-style many-pattern union DFA (built here by the
compiled reference: per pattern re_comp/determinise/minimise/setendid, fsm_union_array,
determinise; reference src/rx/main.c:487-566,1353,1371) over ragged log lines.  Every record
AND the end-id set of every match are compared with the reference's own fsm_exec walk."""
import numpy as np
import pytest

import reflib
import libfsm_b200 as L

pytestmark = [pytest.mark.gpu]


def build_patterns(npat, rng):
    words = ["ERROR", "WARN", "INFO", "DEBUG", "FATAL", "kernel", "sshd", "nginx", "cron", "systemd"]
    templates = [lambda w, k: (f"^{w} [0-9]{{{k}}} ", f"{w} {'7' * k} "),
                 lambda w, k: (f"^{w}: user=[a-z]+ id=[0-9]{{{k}}}", f"{w}: user=bob id={'4' * k}"),
                 lambda w, k: (f"^{w}\\[[0-9]+\\]: ", f"{w}[123]: "),
                 lambda w, k: (f"^{w} (GET|POST|PUT) /[a-z/]+ ", f"{w} GET /a/b "),
                 lambda w, k: (f"^{w} [A-Z]{{{k}}}-[0-9]+", f"{w} {'Q' * k}-99")]
    pats, exs = [], []
    while len(pats) < npat:
        w = words[int(rng.integers(len(words)))] + str(int(rng.integers(0, 30)))
        p, e = templates[int(rng.integers(len(templates)))](w, int(rng.integers(1, 4)))
        if p not in pats:
            pats.append(p); exs.append(e.encode())
    return pats, exs


@pytest.mark.parametrize("npat", [24, 96])
def test_union_dfa_over_ragged_lines(ref, oracle, npat):
    rng = np.random.default_rng(7 + npat)
    pats, exs = build_patterns(npat, rng)
    h = ref.union_dfa(pats, state_limit=200000)
    fsm = ref.flatten(h)
    nlines = 300000
    lens = rng.integers(0, 257, size=nlines)
    lines = []
    noise = rng.integers(0x20, 0x7F, size=int(lens.sum()) + 64, dtype=np.uint8)
    pos = 0
    for i in range(nlines):
        body = noise[pos:pos + int(lens[i])].tobytes(); pos += int(lens[i])
        if i % 2 == 0:
            e = exs[int(rng.integers(len(exs)))]
            body = (e + body)[:max(len(body), 0)] if i % 4 == 0 else e[:len(e) // 2] + body
        lines.append(body)
    base, off = reflib.offsets_for(lines)
    want = ref.exec_batch(h, base, off, mode=1, nthreads=16)      # the reference's own walk
    with L.Dfa(fsm) as dfa:
        got = dfa.exec_batch(base, off)                            # host path -> ragged kernel
        import torch
        dout = dfa.exec_batch(torch.from_numpy(base).cuda(), torch.from_numpy(off.astype(np.int64)).cuda())
        torch.cuda.synchronize()
        assert dfa.info["entry_bytes"] == 2 or fsm.nstates <= 255
    assert (got == want).all()
    assert (L.results_from_torch(dout) == want).all()
    assert 0.05 < (want["ret"] == 1).mean() < 0.9
    # end ids through the flat description == fsm_endid_get of the reference
    for i in np.nonzero(want["ret"] == 1)[0][:500]:
        assert list(fsm.endids_of(int(got["end"][i]))) == ref.endids(h, int(want["end"][i]))
    ref.free(h)
