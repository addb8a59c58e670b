"""The oracle (oracle/lf_oracle.c) against the committed golden vectors (generated from the
reference's own edlib / ksw / alignChain_edlib by tests/golden/make_golden.py) and, where the
reference binaries were built in this container, against the reference directly."""
import random

import numpy as np
import pytest

import _oracle as O
from _common import CODE, load_chains, load_ksw, load_pairs
from lordfast_b200 import sim


def test_golden_align_pairs():
    pairs = load_pairs()
    assert len(pairs) > 200
    for p in pairs:
        ed, end, ops = O.oracle_align(p["q"].encode(), p["t"].encode(), p["mode"])
        assert (ed, end) == (p["ed"], p["end"])
        assert "".join(str(c) for c in ops) == p["ops"]


def test_golden_ksw():
    for c in load_ksw():
        q = CODE[np.frombuffer(c["q"].encode(), dtype=np.uint8)].tobytes()
        t = CODE[np.frombuffer(c["t"].encode(), dtype=np.uint8)].tobytes()
        assert O.oracle_extend(q, t, *c["prm"]) == (c["score"], c["qle"], c["tle"])


def test_golden_chains():
    z, chains = load_chains()
    l_pac = int(z["l_pac"])
    ref = sim.ACGT[np.array([O.oracle().lfo_pac_get(z["pac"].ctypes.data, i) for i in range(0)], dtype=np.uint8)] if False else None
    # unpack the 2-bit reference written by the reference's own `--index`
    pac = z["pac"]
    idx = np.arange(l_pac)
    ref = sim.ACGT[(pac[idx >> 2] >> ((~idx & 3) << 1)) & 3]
    ridx = O.RefIndex(ref.tobytes())
    assert bytes(ridx.pac.raw[: l_pac // 4]) == pac[: l_pac // 4].tobytes()
    reads, off = z["reads"], z["read_off"]
    nsplit = 0
    for c in chains:
        r = reads[off[c["read"]]:off[c["read"] + 1]]
        q = (sim.revcomp(r) if c["isRev"] else r).tobytes()
        got, st = O.oracle_align_chain(ridx, [tuple(s) for s in c["seeds"]], q, c["isRev"])
        assert got == c["sam"]
        nsplit += len(c["sam"]) > 1
    assert nsplit >= 3  # the fixtures exercise the split / clip paths


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_vs_reference_random():
    rng = random.Random(7)

    def rnd(n):
        return bytes(rng.choice(b"ACGT") for _ in range(n))

    def mut(s, e):
        out = bytearray()
        for ch in s:
            r = rng.random()
            if r < e * 0.1:
                out.append(rng.choice(b"ACGT"))
            elif r < e * 0.7:
                out.append(ch); out.append(rng.choice(b"ACGT"))
            elif r < e:
                pass
            else:
                out.append(ch)
        return bytes(out) or b"A"

    for it in range(1500):
        L = rng.choice([1, 2, 3, 5, 8, 13, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 200, 300, 520])
        t = rnd(L)
        k = rng.random()
        q = mut(t, rng.choice([0.05, 0.15, 0.3])) if k < 0.6 else rnd(rng.randint(1, 2 * L + 3)) if k < 0.8 else mut(t[: rng.randint(1, L)], 0.15)
        for mode in (0, 1):
            tt = t + (rnd(20) if mode == 1 else b"")
            assert O.oracle_align(q, tt, mode) == O.ref_align(q, tt, mode)
    for it in range(8):
        L = rng.choice([1793, 2100, 3000])
        t = rnd(L)
        q = mut(t, 0.15) if it % 2 == 0 else rnd(rng.randint(40, 2500))
        for mode in (0, 1):
            assert O.oracle_align(q, t, mode) == O.ref_align(q, t, mode)
    code = lambda s: bytes(b"ACGT".index(c) for c in s)
    for it in range(600):
        L = rng.choice([5, 50, 200, 600, 1200])
        t = rnd(L)
        k = rng.random()
        q = mut(t[: L // 2], 0.15) + rnd(L // 2) if k < 0.4 else mut(t, 0.2) if k < 0.7 else rnd(rng.randint(1, L + 50))
        for prm in ((0, 1, 0, 1, 40, 40), (8, 1, 4, 1, 100, 200)):
            assert O.oracle_extend(code(q), code(t), *prm) == O.ref_extend(code(q), code(t), *prm)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_chain_vs_reference():
    w = sim.make_workload(400_000, 60, 5000, 0.12, 0.15, seed=3, sv_frac=0.5, sv_kinds=sim.SV_KINDS + ("inversion_del",))
    idx = O.RefIndex(w.ref.tobytes())
    nex = 0
    ninv = 0
    for i in range(w.n_reads):
        seeds = [tuple(int(x) for x in s) for s in w.chain(i)]
        q = w.oriented(i).tobytes()
        a, st = O.oracle_align_chain(idx, seeds, q, int(w.is_rev[i]))
        assert a == O.ref_align_chain(idx, seeds, q, int(w.is_rev[i]))
        nex += st.n_extend
        ninv += any((r["flag"] & 16) != (16 if w.is_rev[i] else 0) for r in a)
    assert nex > 0 and ninv > 0  # clip / split rounds and the accepted-inversion branch are covered
