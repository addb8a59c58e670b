"""lf_gpu_align_chains (the batched alignChain_edlib) against the Sam_t records the reference itself
produced for the golden chains, and against the oracle's chain restatement on simulated SV reads.
CPU runs go through the test-only emulator; the -m gpu runs through liblfgpu.so on the device."""
import numpy as np
import pytest

import _oracle as O
from _common import CRAFTED_REF, build_emu, load_chains, load_crafted
from lordfast_b200 import api, sim


def _golden(lib_path):
    z, chains = load_chains()
    l_pac = int(z["l_pac"])
    g = api.LfGpu(z["pac"], l_pac, lib_path=lib_path)
    seeds, ch = [], []
    for c in chains:
        ch.append((len(seeds), len(c["seeds"]), c["read"], c["isRev"], 0))
        seeds.extend(tuple(s) for s in c["seeds"])
    recs, text, st = g.align_chains(z["reads"], z["read_off"].astype(np.uint64), [0], [l_pac],
                                    np.array(seeds, dtype=api.SEED), np.array(ch, dtype=api.CHAIN))
    got = api.records_to_dicts(recs, text)
    exp = []
    for ci, c in enumerate(chains):
        for s in c["sam"]:
            d = dict(chain=ci); d.update(s); exp.append(d)
    assert got == exp
    assert st.round2_extends > 0 and st.round3_tasks > 0  # the clip / split rounds really ran
    g.close()


def _simulated(lib_path, n_reads, read_len, ref_len, need_inversion=True, sv_frac=0.6):
    w = sim.make_workload(ref_len, n_reads, read_len, 0.12, 0.15, seed=21, sv_frac=sv_frac, sv_kinds=sim.SV_KINDS + ("inversion_del", "inversion_del"))
    g = api.LfGpu(w.pac, len(w.ref), lib_path=lib_path)
    seeds, chains = api.workload_chains(w)
    recs, text, st = g.align_chains(w.reads, w.read_off.astype(np.uint64), w.contig_off, w.contig_len, seeds, chains)
    got = api.records_to_dicts(recs, text)
    idx = O.RefIndex(w.ref.tobytes())
    exp = []
    for i in range(w.n_reads):
        a, _ = O.oracle_align_chain(idx, [tuple(int(x) for x in s) for s in w.chain(i)], w.oriented(i).tobytes(), int(w.is_rev[i]))
        for s in a:
            d = dict(chain=i); d.update(s); exp.append(d)
    assert got == exp
    # the accepted-inversion branch (MD / CIGAR out of step, src/LordFAST.cpp:2056-2057) is exercised
    assert not need_inversion or any((r["flag"] & 16) != (16 if w.is_rev[r["chain"]] else 0) for r in exp)
    g.close()
    return st


def _crafted(lib_path):
    """Hand-built chains with the reference's recorded answers (tests/golden/make_crafted_chains.py): records
    dropped for having fewer than two anchors, splits in adjacent gaps, error-free stretches."""
    ref = sim.make_reference(*CRAFTED_REF)
    cases = load_crafted()
    g = api.LfGpu(sim.pack_pac(ref), len(ref), lib_path=lib_path)
    reads = [np.frombuffer(c["read"].encode(), dtype=np.uint8) for c in cases]
    read_off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
    seeds, ch = [], []
    for i, c in enumerate(cases):
        ch.append((len(seeds), len(c["seeds"]), i, 0, 0))
        seeds.extend(tuple(s) for s in c["seeds"])
    recs, text, st = g.align_chains(np.concatenate(reads), read_off, [0], [len(ref)], np.array(seeds, dtype=api.SEED), np.array(ch, dtype=api.CHAIN))
    got = api.records_to_dicts(recs, text)
    exp = []
    for ci, c in enumerate(cases):
        for s in c["sam"]:
            d = dict(chain=ci); d.update(s); exp.append(d)
    assert got == exp
    assert any(len(c["sam"]) == 2 and len(c["seeds"]) > 50 for c in cases)  # a dropped middle record is in the set
    g.close()


def _long_ends(lib_path, n_reads=10, read_len=4000):
    """Heads / tails longer than _pf_clipLen that are NOT junk: round 1 only measures them (distance), the clip test
    fails (similarity >= 0.75) or the extension keeps all of them, and round 3 has to deliver the prefix-mode path
    (src/LordFAST.cpp:1869-1886, :2196-2217).  Made by dropping the outer anchors of simulated chains."""
    w = sim.make_workload(150_000, n_reads, read_len, 0.05, 0.15, seed=33, sv_frac=0.3, sv_kinds=("junk_head", "junk_tail"))
    seeds, chains = api.workload_chains(w)
    keep, new_chains = [], []
    for i in range(w.n_reads):
        lo, hi = int(w.seed_off[i]), int(w.seed_off[i + 1])
        cut = (7, 7) if i % 3 == 0 else (9, 0) if i % 3 == 1 else (0, 12)
        if hi - lo - cut[0] - cut[1] < 2:
            cut = (0, 0)
        new_chains.append((len(keep), hi - lo - cut[0] - cut[1], i, int(w.is_rev[i]), 0))
        keep.extend(range(lo + cut[0], hi - cut[1]))
    seeds = seeds[np.array(keep)]
    chains = np.array(new_chains, dtype=api.CHAIN)
    g = api.LfGpu(w.pac, len(w.ref), lib_path=lib_path)
    recs, text, st = g.align_chains(w.reads, w.read_off.astype(np.uint64), w.contig_off, w.contig_len, seeds, chains)
    got = api.records_to_dicts(recs, text)
    idx = O.RefIndex(w.ref.tobytes())
    exp, long_kept = [], 0
    for i, c in enumerate(chains):
        sd = [(int(x["tPos"]), int(x["qPos"]), int(x["len"])) for x in seeds[int(c["seed_off"]):int(c["seed_off"]) + int(c["n_seeds"])]]
        a, _ = O.oracle_align_chain(idx, sd, w.oriented(i).tobytes(), int(w.is_rev[i]))
        for r in a:
            d = dict(chain=i); d.update(r); exp.append(d)
        long_kept += sum(1 for r in a if sd[0][1] > 500 and r["qStart"] == 0)
    assert got == exp
    assert long_kept > 0 and st.round3_tasks > 0   # a long head was kept whole: its path came from round 3
    g.close()


def test_emu_chain_operator_long_ends(monkeypatch):
    _long_ends(build_emu())
    monkeypatch.setenv("LF_CHAIN_HOST_EMIT", "1")
    _long_ends(build_emu())


@pytest.mark.gpu
def test_gpu_chain_operator_long_ends(monkeypatch):
    _long_ends(None, 60, 8000)
    monkeypatch.setenv("LF_CHAIN_NO_SPEC", "1")
    _long_ends(None, 60, 8000)


def test_oracle_crafted_chains():
    ref = sim.make_reference(*CRAFTED_REF)
    idx = O.RefIndex(ref.tobytes())
    for c in load_crafted():
        a, _ = O.oracle_align_chain(idx, [tuple(s) for s in c["seeds"]], c["read"].encode(), 0)
        assert a == c["sam"]


def test_emu_chain_operator_crafted(monkeypatch):
    _crafted(build_emu())
    monkeypatch.setenv("LF_CHAIN_HOST_EMIT", "1")
    _crafted(build_emu())


@pytest.mark.gpu
def test_gpu_chain_operator_crafted(monkeypatch):
    _crafted(None)
    monkeypatch.setenv("LF_CHAIN_HOST_EMIT", "1")
    _crafted(None)


def test_emu_chain_operator_long_reads():
    """More anchors than one block has threads: the slot carries cross rounds."""
    _simulated(build_emu(), 4, 14_000, 300_000, need_inversion=False)


def test_emu_chain_operator_golden():
    _golden(build_emu())


def test_emu_chain_operator_simulated():
    _simulated(build_emu(), 14, 3000, 120_000)


@pytest.mark.parametrize("lanes,slow", [("3", "1"), ("2", "0")])
def test_emu_chain_operator_lanes(monkeypatch, lanes, slow):
    """One call cut into pipelined lanes (lf_chain.inl): the chains that can reach rounds 2 / 3 or hold long tasks go to the
    slow lane, whose reads are gathered and uploaded first; the others to consecutive fast lanes.  Records come back in
    chain order whatever the split."""
    monkeypatch.setenv("LF_CHAIN_LANES", lanes)
    monkeypatch.setenv("LF_CHAIN_SLOW_LANE", slow)
    _simulated(build_emu(), 24, 3000, 200_000, need_inversion=False, sv_frac=0.2)
    _golden(build_emu())


@pytest.mark.gpu
@pytest.mark.parametrize("lanes,slow", [("4", "1"), ("3", "0"), ("1", "1")])
def test_gpu_chain_operator_lanes(monkeypatch, lanes, slow):
    monkeypatch.setenv("LF_CHAIN_LANES", lanes)
    monkeypatch.setenv("LF_CHAIN_SLOW_LANE", slow)
    _simulated(None, 400, 8000, 1_500_000, sv_frac=0.15)
    _golden(None)


def test_emu_chain_operator_host_emit(monkeypatch):
    """The alternative emit path (host threads from the 2-bit op stream, used by multi-device contexts)."""
    monkeypatch.setenv("LF_CHAIN_HOST_EMIT", "1")
    _golden(build_emu())
    _simulated(build_emu(), 14, 3000, 120_000)


@pytest.mark.gpu
def test_gpu_chain_operator_host_emit(monkeypatch):
    monkeypatch.setenv("LF_CHAIN_HOST_EMIT", "1")
    _golden(None)
    _simulated(None, 300, 6_000, 1_000_000)


@pytest.mark.gpu
def test_gpu_chain_operator_golden():
    _golden(None)


@pytest.mark.gpu
def test_gpu_chain_operator_simulated_config1():
    st = _simulated(None, 200, 10_000, 1_000_000)
    assert st.round2_extends > 0


@pytest.mark.gpu
def test_gpu_chain_operator_two_contigs_edges():
    """Chains next to contig boundaries: the head / tail guards (:1825, :2163) must soft-clip."""
    rng = np.random.default_rng(3)
    ref = sim.make_reference(60_000, 8)
    contig_off, contig_len = [0, 25_000], [25_000, 35_000]
    reads, seeds, chains = [], [], []
    idx = O.RefIndex(ref.tobytes(), contig_len)
    exp = []
    for k, (start, L) in enumerate([(5, 3000), (25_010, 3000), (21_990, 3000), (56_990, 3000), (40_000, 2000)]):
        o, qpos, clean = sim._channel(ref[start:start + L], 0.12, rng)
        s = sim._anchors(qpos, clean, start, 0, 14, rng)
        if k % 2:
            o = np.concatenate([sim.ACGT[rng.integers(0, 4, size=40, dtype=np.uint8)], o]); s[:, 1] += 40
        rev = k % 2
        reads.append(sim.revcomp(o) if rev else o)
        chains.append((len(seeds), len(s), k, rev, 0))
        seeds.extend(tuple(int(x) for x in r) for r in s)
        a, _ = O.oracle_align_chain(idx, [tuple(int(x) for x in r) for r in s], o.tobytes(), rev)
        for r in a:
            d = dict(chain=k); d.update(r); exp.append(d)
    off = np.zeros(len(reads) + 1, dtype=np.uint64); off[1:] = np.cumsum([len(r) for r in reads])
    g = api.LfGpu(sim.pack_pac(ref), len(ref))
    recs, text, st = g.align_chains(np.concatenate(reads), off, contig_off, contig_len, np.array(seeds, dtype=api.SEED), np.array(chains, dtype=api.CHAIN))
    assert api.records_to_dicts(recs, text) == exp
    g.close()


@pytest.mark.parametrize("env", ["LF_CHAIN_NO_SPEC", "LF_CHAIN_HOST_TASKS", "LF_CHAIN_HOST_PLAN", "LF_EMIT_NO_ZEROCOPY"])
def test_emu_chain_operator_fallback_paths(monkeypatch, env):
    """The non-default variants of the single-device path: round 2 as its own GPU round trip instead of the speculative
    extensions, round-1 tasks built on host threads instead of by k_chain_tasks, emit text through a staging copy."""
    monkeypatch.setenv(env, "1")
    _crafted(build_emu())
    _golden(build_emu())


@pytest.mark.gpu
@pytest.mark.parametrize("env", ["LF_CHAIN_NO_SPEC", "LF_CHAIN_HOST_TASKS", "LF_CHAIN_HOST_PLAN", "LF_EMIT_NO_ZEROCOPY"])
def test_gpu_chain_operator_fallback_paths(monkeypatch, env):
    monkeypatch.setenv(env, "1")
    _crafted(None)
    _golden(None)
    _simulated(None, 200, 6_000, 1_000_000)


def _edge_cases(lib_path):
    """Empty chain list, a chain without any gap / head / tail task (one merged match run), a chain naming a read that does
    not exist (rejected with LF_ERR_BAD_ARG, nothing aligned)."""
    ref = sim.make_reference(50_000, 3)
    g = api.LfGpu(sim.pack_pac(ref), len(ref), lib_path=lib_path)
    reads = ref[1000:3000].copy()
    off = np.array([0, 2000], dtype=np.uint64)
    recs, text, st = g.align_chains(reads, off, [0], [len(ref)], np.zeros(0, dtype=api.SEED), np.zeros(0, dtype=api.CHAIN))
    assert len(recs) == 0 and st.round1_tasks == 0
    seeds = np.array([(1000, 0, 1000), (2000, 1000, 1000)], dtype=api.SEED)
    recs, text, st = g.align_chains(reads, off, [0], [len(ref)], seeds, np.array([(0, 2, 0, 0, 0)], dtype=api.CHAIN))
    got = api.records_to_dicts(recs, text)
    idx = O.RefIndex(ref.tobytes())
    exp, _ = O.oracle_align_chain(idx, [(1000, 0, 1000), (2000, 1000, 1000)], reads.tobytes(), 0)
    assert [{k: v for k, v in r.items() if k != "chain"} for r in got] == exp
    with pytest.raises(api.LfGpuError):
        g.align_chains(reads, off, [0], [len(ref)], seeds, np.array([(0, 2, 5, 0, 0)], dtype=api.CHAIN))
    # chains the reference's chaining cannot produce are rejected, not aligned: overlapping seeds, seeds out of order, a seed past
    # the end of the read, a seed past the end of the reference (k_chain_plan, or the host loop with LF_CHAIN_HOST_PLAN=1)
    for bad in ([(1000, 0, 1000), (1900, 900, 1000)], [(2000, 1000, 500), (1000, 0, 500)], [(1000, 0, 1000), (2000, 1000, 1001)],
                [(1000, 0, 1000), (len(ref) - 500, 1000, 1000)]):
        with pytest.raises(api.LfGpuError):
            g.align_chains(reads, off, [0], [len(ref)], np.array(bad, dtype=api.SEED), np.array([(0, 2, 0, 0, 0)], dtype=api.CHAIN))
    recs, text, st = g.align_chains(reads, off, [0], [len(ref)], seeds, np.array([(0, 2, 0, 0, 0)], dtype=api.CHAIN))   # and the context is still usable
    assert len(recs) == 1
    g.close()


def test_emu_chain_operator_edge_cases_host_plan(monkeypatch):
    monkeypatch.setenv("LF_CHAIN_HOST_PLAN", "1")
    _edge_cases(build_emu())


def test_emu_chain_operator_edge_cases():
    _edge_cases(build_emu())


@pytest.mark.gpu
def test_gpu_chain_operator_edge_cases():
    _edge_cases(None)


def _resident_reads(lib_path):
    """reads->bases == NULL: lf_gpu_align_chains works on the reads an earlier call left on the device (what the drop-in
    program does after seeding a chunk); a different read count is refused."""
    w = sim.make_workload(120_000, 10, 2500, 0.12, 0.15, seed=8, sv_frac=0.5)
    g = api.LfGpu(w.pac, len(w.ref), lib_path=lib_path)
    seeds, chains = api.workload_chains(w)
    off = w.read_off.astype(np.uint64)
    want = g.align_chains(w.reads, off, w.contig_off, w.contig_len, seeds, chains)
    g.upload_reads(w.reads, off)
    got = g.align_chains(None, off, w.contig_off, w.contig_len, seeds, chains)
    assert np.array_equal(got[0], want[0]) and got[1] == want[1] and len(want[0]) > 0
    with pytest.raises(api.LfGpuError):
        g.align_chains(None, off[:-1], w.contig_off, w.contig_len, seeds, chains[chains["read_id"] < len(off) - 2])
    g.close()


def test_emu_chain_operator_resident_reads():
    _resident_reads(build_emu())


@pytest.mark.gpu
def test_gpu_chain_operator_resident_reads():
    _resident_reads(None)
