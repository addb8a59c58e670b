"""The chain operator on chains the REFERENCE's own front-end produced (lordfast_b200/fixtures.py, written by
tools/make_fixtures.py from oracle/_ref/lordfast_chaindump): every Sam_t record the reference's alignChain_edlib pushed
for them -- flag, pos, posEnd, qStart, qEnd, NM, and the CIGAR / MD text by length + CRC-32 -- against
lf_gpu_align_chains.  Small config-3 / config-4 shaped sets are committed (tests/golden/chains_mini*.npz); the
full-size ones of BASELINE configs[1..3] are built by __graft_entry__.build() into fixtures/ and travel to the GPU box."""
import zlib

import numpy as np
import pytest

import _oracle as O
from _common import build_emu
from lordfast_b200 import api, fixtures, sim


def _run(fx, lib_path=None, t_shift=0, pac=None, l_pac=None, contigs=None, chunk=None):
    g = api.LfGpu(fx.pac if pac is None else pac, len(fx.ref) if l_pac is None else l_pac, lib_path=lib_path)
    co, cl = (fx.contig_off, fx.contig_len) if contigs is None else contigs
    n = fx.n_reads
    step = chunk or n
    bad, nrec, st_sum = [], 0, [0, 0, 0]
    for lo in range(0, n, step):
        sub = fx.subset(lo, min(n, lo + step))
        seeds = sub["seeds"].copy()
        seeds["tPos"] += np.uint32(t_shift)
        recs, text, st = g.align_chains(sub["reads"], sub["read_off"], co, cl, seeds, sub["chains"])
        bad += fx.compare(recs, text, t_shift=t_shift, chain_ids=sub["chain_ids"])
        nrec += len(recs)
        st_sum = [st_sum[0] + st.round1_tasks, st_sum[1] + st.round2_extends, st_sum[2] + st.round3_tasks]
    g.close()
    return bad, nrec, st_sum


@pytest.mark.parametrize("name", ["mini3", "mini4"])
def test_oracle_restatement_on_reference_chains(name):
    """pins oracle/lf_oracle.c's alignChain restatement to the reference's records on config-3 / config-4 shaped chains"""
    fx = fixtures.load(name)
    idx = O.RefIndex(fx.ref.tobytes())
    k = 0
    for ci, ch in enumerate(fx.chains):
        sd = fx.seeds[int(ch["seed_off"]):int(ch["seed_off"]) + int(ch["n_seeds"])]
        rid = int(ch["read_id"])
        r = fx.w.reads[fx.w.read_off[rid]:fx.w.read_off[rid + 1]]
        q = (sim.revcomp(r) if ch["is_rev"] else r).tobytes()
        out, _ = O.oracle_align_chain(idx, [(int(s["tPos"]), int(s["qPos"]), int(s["len"])) for s in sd], q, int(ch["is_rev"]))
        for r in out:
            j = k; k += 1
            assert int(fx.rec["chain"][j]) == ci
            assert (r["flag"], r["pos"], r["posEnd"], r["qStart"], r["qEnd"], r["nm"]) == tuple(int(fx.rec[f][j]) for f in ("flag", "pos", "posEnd", "qStart", "qEnd", "nm"))
            assert (zlib.crc32(r["cigar"].encode()), zlib.crc32(r["md"].encode())) == (int(fx.rec["cigar_crc"][j]), int(fx.rec["md_crc"][j]))
        if str(ci) in fx.full:
            assert [(r["cigar"], r["md"]) for r in out] == [(f["cigar"], f["md"]) for f in fx.full[str(ci)]]
    assert k == len(fx.rec["chain"])
    if name == "mini4":
        assert len(fx.chains) > fx.params["reads_mapped"]   # the fine mode aligned several candidate chains per read


@pytest.mark.parametrize("name,lanes", [("mini3", "1"), ("mini4", "1"), ("mini4", "3")])
def test_emu_reference_chains(name, lanes, monkeypatch):
    """lanes: the call cut into that many pipelined sub-batches (on the emulator they run one after the other)"""
    monkeypatch.setenv("LF_CHAIN_LANES", lanes)
    fx = fixtures.load(name)
    bad, nrec, st = _run(fx, build_emu())
    assert not bad, bad
    assert nrec == len(fx.rec["chain"]) and st[0] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mini3", "mini4"])
def test_gpu_reference_chains_small(name, monkeypatch):
    fx = fixtures.load(name)
    bad, nrec, st = _run(fx)
    assert not bad, bad
    bad, _, _ = _run(fx, chunk=17)   # the same reads in ragged chunks
    assert not bad, bad
    for lanes in ("1", "2", "7"):    # and as pipelined sub-batches of one call
        monkeypatch.setenv("LF_CHAIN_LANES", lanes)
        bad, nrec, _ = _run(fx)
        assert not bad and nrec == len(fx.rec["chain"]), (lanes, bad)


@pytest.mark.gpu
@pytest.mark.parametrize("name,chunk", [("config2", 20_000), ("config3", 5_000), ("config4", 2_500)])
def test_gpu_reference_chains_full(name, chunk):
    """BASELINE configs[1], [2] and a configs[3]-shaped set: every record of every chain against the reference's."""
    if not fixtures.available(name):
        pytest.skip(f"fixtures/{name}.npz not built (tools/make_fixtures.py needs the reference; build() writes it)")
    fx = fixtures.load(name)
    bad, nrec, st = _run(fx, chunk=chunk)
    assert not bad, bad
    assert nrec == len(fx.rec["chain"])
    assert st[1] > 0 and st[2] > 0   # the clip / split rounds ran


@pytest.mark.gpu
def test_gpu_reference_chains_above_2g():
    """configs[3]: a 3.1 Gbp 2-bit reference (775 MB, replicated per GPU).  The indexed 256 Mbp reference (or the small
    config-4 set when the full one is not built) sits at 2.8 Gbp inside it, so every reference offset is above 2^31;
    the expected records are the reference's own with pos / posEnd moved by the same amount, and the reference's
    alignChain_edlib is re-run on a sample of the moved chains when oracle/_ref is present."""
    name = "config4" if fixtures.available("config4") else "mini4"
    fx = fixtures.load(name)
    shift, L = fixtures.T_SHIFT_CONFIG4, fixtures.L_PAC_HUMAN
    pac = np.zeros(L // 4 + 1, dtype=np.uint8)
    pac[shift // 4: shift // 4 + len(fx.pac) - 1] = fx.pac[:-1]   # the shift is a multiple of 4 bases; the last byte is padding
    n = len(fx.ref)
    tail = L - shift - n
    co = np.array([0, shift // 2, shift, shift + n], dtype=np.int64)
    cl = np.array([shift // 2, shift // 2, n, tail], dtype=np.int32)
    bad, nrec, st = _run(fx, t_shift=shift, pac=pac, l_pac=L, contigs=(co, cl), chunk=2_500)
    assert not bad, bad
    assert nrec == len(fx.rec["chain"])
    if O.have_ref():
        idx = O.PacIndex(pac, L, co, cl)
        for ci in range(0, len(fx.chains), max(1, len(fx.chains) // 40)):
            ch = fx.chains[ci]
            rid = int(ch["read_id"])
            r = fx.w.reads[fx.w.read_off[rid]:fx.w.read_off[rid + 1]]
            q = (sim.revcomp(r) if ch["is_rev"] else r).tobytes()
            sd = [(int(s["tPos"]) + shift, int(s["qPos"]), int(s["len"])) for s in fx.seeds[int(ch["seed_off"]):int(ch["seed_off"]) + int(ch["n_seeds"])]]
            out = O.ref_align_chain(idx, sd, q, int(ch["is_rev"]))
            js = np.flatnonzero(fx.rec["chain"] == ci)
            assert [(o["pos"], o["posEnd"], o["nm"]) for o in out] == [(int(fx.rec["pos"][j]) + shift, int(fx.rec["posEnd"][j]) + shift, int(fx.rec["nm"][j])) for j in js]
