"""The kernel + pipeline SOURCE run on the test-only fiber emulator (tests/emu) against the oracle.
This is a debugging aid for a container without a GPU; the parity tests proper are test_gpu_parity.py."""
import numpy as np
import pytest

from _common import CODE, build_emu, check_align, group_class_batch, load_ksw, load_pairs, pairs_as_batch
import _oracle as O
from lordfast_b200 import api, sim
from lordfast_b200.chain_tasks import workload_tasks


@pytest.fixture(scope="module")
def emu_lib():
    return build_emu()


def test_emu_golden_pairs(emu_lib):
    pairs = [p for p in load_pairs() if set(p["t"]) <= set("ACGT") and len(p["q"]) * len(p["t"]) < 1_500_000]
    ref, reads, tasks = pairs_as_batch(pairs)
    g = api.LfGpu(sim.pack_pac(ref), len(ref), lib_path=emu_lib)
    bad, res, ops = check_align(g, reads, ref, tasks)
    assert not bad
    for i, p in enumerate(pairs):  # and against the reference's recorded answers
        assert (int(res[i]["edit_distance"]), int(res[i]["end_location"])) == (p["ed"], p["end"])
        assert "".join(str(c) for c in api.decode_ops(ops, int(res[i]["ops_off"]), int(res[i]["ops_len"]))) == p["ops"]
    g.close()


def test_emu_random_tasks_all_flags(emu_lib):
    rng = np.random.default_rng(5)
    ref = sim.make_reference(30000, 3)
    g = api.LfGpu(sim.pack_pac(ref), len(ref), lib_path=emu_lib)
    reads = []
    for i in range(8):
        L = int(rng.integers(300, 1500)); s = int(rng.integers(100, len(ref) - L - 100))
        reads.append(sim._channel(ref[s:s + L], 0.15, rng)[0])
    r0 = reads[0].copy(); r0[5] = ord("N"); r0[17] = ord("a"); reads[0] = r0
    tasks = []
    qlens = [1, 2, 31, 32, 33, 64, 65, 96, 97, 129, 200, 257, 385, 513, 600]
    for k in range(150):
        rid = int(rng.integers(0, len(reads))); L = len(reads[rid])
        ql = int(min(qlens[k % len(qlens)], L)); qo = int(rng.integers(0, L - ql + 1))
        tl = max(1, int(ql * rng.uniform(0.5, 1.5)) + int(rng.integers(-3, 4)))
        to = int(rng.integers(0, len(ref) - tl))
        flags = int(rng.choice([0, 1, 2, 3, 4, 5])) | (8 if k % 23 == 0 else 0)
        tasks.append((rid, qo, ql, to, tl, flags, int(rng.integers(0, 2)), 0))
    tasks = np.array(tasks, dtype=api.ALIGN_TASK)
    bad, _, _ = check_align(g, reads, ref, tasks)
    assert not bad
    g.close()


def test_emu_bad_tasks_are_flagged(emu_lib):
    ref = sim.make_reference(2000, 1)
    g = api.LfGpu(sim.pack_pac(ref), len(ref), lib_path=emu_lib)
    reads = [ref[100:200].copy()]
    tasks = np.array([(0, 0, 50, 0, 60, 0, 0, 0), (0, 90, 50, 0, 60, 0, 0, 0), (0, 0, 0, 0, 60, 0, 0, 0), (0, 0, 10, 1990, 60, 0, 0, 0), (3, 0, 10, 0, 60, 0, 0, 0)],
                     dtype=api.ALIGN_TASK)
    offs = np.array([0, 100], dtype=np.uint64)
    res, ops = g.align_batch(reads[0], offs, tasks)
    assert int(res[0]["status"]) == 0
    assert [int(s) for s in res["status"][1:]] == [-3, -3, -3, -3]
    g.close()


def test_emu_workload_tasks(emu_lib):
    w = sim.make_workload(100_000, 6, 2500, 0.15, 0.15, seed=2, sv_frac=0.5)
    tasks, chain, kind = workload_tasks(w)
    g = api.LfGpu(w.pac, len(w.ref), lib_path=emu_lib)
    reads = [w.reads[w.read_off[i]:w.read_off[i + 1]] for i in range(w.n_reads)]
    bad, _, _ = check_align(g, reads, w.ref, tasks)
    assert not bad
    g.close()


def test_emu_extend(emu_lib):
    cases = [c for c in load_ksw() if len(c["q"]) < 1000][:30]
    tcat, reads, tasks, off = [], [], [], 0
    for k, c in enumerate(cases):
        t = np.frombuffer(c["t"].encode(), dtype=np.uint8)
        tcat.append(t); reads.append(np.frombuffer(c["q"].encode(), dtype=np.uint8))
        p = c["prm"]
        tasks.append((k, 0, len(c["q"]), off, len(t), 0, 0, 0, p[0], p[1], p[2], p[3], p[4], p[5], len(c["q"])))
        off += len(t)
    ref = np.concatenate(tcat)
    g = api.LfGpu(sim.pack_pac(ref), len(ref), lib_path=emu_lib)
    offs = np.zeros(len(reads) + 1, dtype=np.uint64); offs[1:] = np.cumsum([len(r) for r in reads])
    er = g.extend_batch(np.concatenate(reads), offs, np.array(tasks, dtype=api.EXTEND_TASK))
    for k, c in enumerate(cases):
        assert (int(er[k]["score"]), int(er[k]["qle"]), int(er[k]["tle"])) == (c["score"], c["qle"], c["tle"])
    # the same pairs with a band too wide for the shared-memory ring (eh[] in global scratch), against the oracle
    wide = np.array(tasks, dtype=api.EXTEND_TASK)
    wide["w"] = 300
    er = g.extend_batch(np.concatenate(reads), offs, wide)
    for k, c in enumerate(cases):
        p = c["prm"]
        exp = O.oracle_extend(CODE[reads[k]].tobytes(), CODE[tcat[k]].tobytes(), p[0], p[1], p[2], p[3], 300, p[5])
        assert exp == (int(er[k]["score"]), int(er[k]["qle"]), int(er[k]["tle"]))
    g.close()


def test_emu_large_and_hirschberg_tasks(emu_lib):
    """k_myers_large: several words per lane, two strips (q > 8192 rows), SHW with a late best column,
    and the Hirschberg recursion (sizes above edlib's 1 MiB rule)."""
    rng = np.random.default_rng(9)
    ref = sim.make_reference(40000, 4)
    g = api.LfGpu(sim.pack_pac(ref), len(ref), lib_path=emu_lib)
    reads, src = [], []
    for L in (700, 1500, 2600, 9000):
        s = int(rng.integers(100, 20000))
        reads.append(sim._channel(ref[s:s + L], 0.15, rng)[0]); src.append(s)
    tasks = []
    for rid in range(4):
        L = len(reads[rid]); s = src[rid]; tl = int(L / 1.045)
        if rid < 3:
            tasks.append((rid, 0, L, s, tl, 0, 0, 0))
            tasks.append((rid, 0, L, s, tl + 20, 0, 1, 0))
            tasks.append((rid, 0, L, s, tl, 2, 0, 0))
            tasks.append((rid, 0, L, int(rng.integers(0, 10000)), max(2, L // 2), 0, 1, 0))
    tasks.append((3, 0, len(reads[3]), src[3], 300, 0, 0, 0))    # 2 strips of 8 words per lane, leaf
    tasks.append((3, 0, len(reads[3]), src[3], 320, 1, 1, 0))    # same, SHW, reverse strand
    tasks.append((3, 0, len(reads[3]), src[3], 700, 0, 0, 0))    # 2 strips + Hirschberg
    tasks.append((1, 0, 40, 100, 27000, 0, 0, 0))                # one word, very long target, Hirschberg
    tasks = np.array(tasks, dtype=api.ALIGN_TASK)
    bad, _, _ = check_align(g, reads, ref, tasks)
    assert not bad, bad
    g.close()


@pytest.mark.parametrize("groupk,wide", [("7", "2048"), ("7", "0"), ("0", "2048"), ("3", "0")])
def test_emu_group_classes(emu_lib, monkeypatch, groupk, wide):
    """k_myers_group (LANES lanes per task) against the oracle, in the narrow shapes (many tasks) and the wide ones (few tasks:
    LF_GROUP_WIDE is the task count below which a class goes wide); LF_GROUPK=0 routes the same tasks to the older kernels."""
    monkeypatch.setenv("LF_GROUPK", groupk)
    monkeypatch.setenv("LF_GROUP_WIDE", wide)
    ref, reads, tasks = group_class_batch(big=groupk == "7")
    g = api.LfGpu(sim.pack_pac(ref), len(ref), lib_path=emu_lib)
    bad, res, _ = check_align(g, reads, ref, tasks)
    assert not bad, bad
    cc = g.class_counts()
    if groupk == "7":
        assert all(cc.get(k, 0) > 0 for k in ("group_path16", "group_path32", "group_path64", "group_dist32", "group_dist64", "group_dist128", "group_dist256", "large")), cc
    if groupk == "0":
        assert not any(k.startswith("group") for k in cc), cc
    g.close()


def _band_stress_batch(rng, ref, n):
    """Pairs of 129..512 rows whose optimal path bulges off the straight diagonal: block indels that a
    sliding band of 3-5 words sometimes covers and sometimes does not."""
    reads, tasks, pos = [], [], 1000
    for k in range(n):
        L = int(rng.integers(129, 520))
        t = ref[pos:pos + L]
        q = sim.mutate_pair(t, float(rng.choice([0.02, 0.06, 0.10, 0.14])), rng)
        kind, junk = k % 5, lambda m: sim.ACGT[rng.integers(0, 4, size=m, dtype=np.uint8)]
        if kind == 1:
            m = int(rng.integers(10, 60)); at = int(rng.integers(0, len(q)))
            q = np.concatenate([q[:at], junk(m), q[at:]])
        elif kind == 2:
            m = int(rng.integers(10, 60)); at = int(rng.integers(0, max(1, len(q) - m)))
            q = np.concatenate([q[:at], q[at + m:]])
        elif kind == 3:
            m = int(rng.integers(10, 45)); a1 = int(rng.integers(0, len(q) // 3)); a2 = int(rng.integers(2 * len(q) // 3, len(q) - m - 1))
            q = np.concatenate([q[:a1], junk(m), q[a1:a2], q[a2 + m:]])
        elif kind == 4:
            m = int(rng.integers(10, 45)); a1 = int(rng.integers(0, len(q) // 3)); a2 = int(rng.integers(2 * len(q) // 3, len(q) - 1))
            q = np.concatenate([q[:a1], q[a1 + m:a2], junk(m), q[a2:]])
        q = q[:512]
        if len(q) < 1:
            continue
        reads.append(q)
        tasks.append((len(reads) - 1, 0, len(q), pos, L, int(rng.integers(0, 2)) * 2, 0, 0))
        pos += L + 20
    return reads, np.array(tasks, dtype=api.ALIGN_TASK)


def test_emu_banded_classes_certificate_and_retry(emu_lib, monkeypatch):
    """k_myers_band with the sliding band switched on for every class it supports (off by default for q > 128):
    certified tasks and tasks redone full-width must both match the oracle."""
    import ctypes as C
    monkeypatch.setenv("LF_BAND_MASK", "0xffff")
    monkeypatch.setenv("LF_BANDREG", "0")        # the register-band kernel would take these classes otherwise
    monkeypatch.setenv("LF_BANDREG_SMALL", "0")
    rng = np.random.default_rng(4)
    ref = sim.make_reference(200_000, 5)
    reads, tasks = _band_stress_batch(rng, ref, 300)
    g = api.LfGpu(sim.pack_pac(ref), len(ref), lib_path=emu_lib)
    ok0, rt0 = C.c_ulong(), C.c_ulong()
    g.lib.lf_emu_band_counts(C.byref(ok0), C.byref(rt0))
    bad, _, _ = check_align(g, reads, ref, tasks)
    assert not bad
    ok, rt = C.c_ulong(), C.c_ulong()
    g.lib.lf_emu_band_counts(C.byref(ok), C.byref(rt))
    assert ok.value - ok0.value > 50 and rt.value - rt0.value > 20  # both outcomes exercised
    g.close()


def _near_diagonal_batch(rng, ref, n, lo=129, hi=512):
    """Gap-like pairs (query = target through the error channel) of every class k_myers_bandreg serves, all
    strand / direction flags."""
    reads, tasks, pos = [], [], 500
    for k in range(n):
        L = int(rng.integers(lo, hi + 8))
        t = ref[pos:pos + L]
        q = sim.mutate_pair(t, float(rng.choice([0.0, 0.05, 0.12, 0.15, 0.22])), rng)[:hi]
        if len(q) < 1:
            continue
        flags = int(rng.choice([0, api.LF_F_READ_REV, api.LF_F_REVERSE_BOTH, api.LF_F_READ_REV | api.LF_F_REVERSE_BOTH,
                                api.LF_F_RC_QUERY, api.LF_F_NO_PATH]))
        if flags & api.LF_F_READ_REV:
            q = sim.revcomp(q)
        if flags & api.LF_F_REVERSE_BOTH:
            q = q[::-1]  # the task reads both slices right-to-left: lay the stored read out so that the pair stays similar
        if flags & api.LF_F_RC_QUERY:
            q = sim.revcomp(q)
        reads.append(np.ascontiguousarray(q))
        tasks.append((len(reads) - 1, 0, len(q), pos, L, flags, 0, 0))
        pos += L + int(rng.integers(0, 40))
    return reads, np.array(tasks, dtype=api.ALIGN_TASK)


@pytest.mark.parametrize("bandreg", ["0xf", "0"])
def test_emu_bandreg_near_diagonal(emu_lib, monkeypatch, bandreg):
    """k_myers_bandreg (sliding register band, checkpoints + recompute) on near-diagonal global tasks of 129..512
    rows, and the same batch with the kernel switched off (full width): both must match the oracle."""
    import ctypes as C
    monkeypatch.setenv("LF_BANDREG", bandreg)
    rng = np.random.default_rng(21)
    ref = sim.make_reference(300_000, 6)
    reads, tasks = _near_diagonal_batch(rng, ref, 420)
    r2, t2 = _band_stress_batch(rng, ref, 120)
    t2["read_id"] += len(reads)
    reads, tasks = reads + r2, np.concatenate([tasks, t2])
    g = api.LfGpu(sim.pack_pac(ref), len(ref), lib_path=emu_lib)
    ok0, rt0 = C.c_ulong(), C.c_ulong()
    g.lib.lf_emu_band_counts(C.byref(ok0), C.byref(rt0))
    bad, res, _ = check_align(g, reads, ref, tasks)
    assert not bad, bad[:10]
    ok, rt = C.c_ulong(), C.c_ulong()
    g.lib.lf_emu_band_counts(C.byref(ok), C.byref(rt))
    if bandreg == "0xf":
        assert ok.value - ok0.value > 250, (ok.value - ok0.value, rt.value - rt0.value)   # certified in the band
        assert rt.value - rt0.value > 5                                                    # and some redone full width
    else:
        assert ok.value == ok0.value
    g.close()
