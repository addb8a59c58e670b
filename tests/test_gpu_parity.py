"""Parity tests proper: the CUDA path, through the C ABI (liblfgpu.so), against the oracle and the
committed golden vectors.  Bit-exact: everything on this path is integer work."""
import numpy as np
import pytest

import _oracle as O
from _common import CODE, check_align, load_chains, load_ksw, load_pairs, pairs_as_batch, task_strings
from lordfast_b200 import api, sim
from lordfast_b200.chain_tasks import workload_tasks

pytestmark = pytest.mark.gpu


def test_gpu_golden_pairs():
    pairs = [p for p in load_pairs() if set(p["t"]) <= set("ACGT")]
    ref, reads, tasks = pairs_as_batch(pairs)
    g = api.LfGpu(sim.pack_pac(ref), len(ref))
    bad, res, ops = check_align(g, reads, ref, tasks)
    assert not bad
    for i, p in enumerate(pairs):
        assert (int(res[i]["edit_distance"]), int(res[i]["end_location"])) == (p["ed"], p["end"])
        assert "".join(str(c) for c in api.decode_ops(ops, int(res[i]["ops_off"]), int(res[i]["ops_len"]))) == p["ops"]
    g.close()


def test_gpu_random_tasks_all_flags():
    rng = np.random.default_rng(55)
    ref = sim.make_reference(200_000, 3)
    g = api.LfGpu(sim.pack_pac(ref), len(ref))
    reads = []
    for i in range(40):
        L = int(rng.integers(300, 6000)); s = int(rng.integers(100, len(ref) - L - 100))
        reads.append(sim._channel(ref[s:s + L], 0.15, rng)[0])
    r0 = reads[0].copy(); r0[5] = ord("N"); r0[17] = ord("a"); r0[100:140] = ord("N"); reads[0] = r0
    tasks = []
    qlens = [1, 2, 5, 31, 32, 33, 63, 64, 65, 90, 96, 97, 128, 129, 190, 200, 256, 257, 300, 384, 385, 500, 512, 513, 600, 700, 1024, 1025, 1500, 2100, 3000]
    for k in range(3000):
        rid = int(rng.integers(0, len(reads))); L = len(reads[rid])
        ql = int(min(qlens[k % len(qlens)], L)); qo = int(rng.integers(0, L - ql + 1))
        tl = max(1, int(ql * rng.uniform(0.5, 1.5)) + int(rng.integers(-3, 4)))
        if k % 17 == 0:
            tl = int(rng.integers(1, 40))
        to = int(rng.integers(0, len(ref) - tl))
        flags = int(rng.choice([0, 1, 2, 3, 4, 5])) | (8 if k % 23 == 0 else 0)
        tasks.append((rid, qo, ql, to, tl, flags, int(rng.integers(0, 2)), 0))
    tasks = np.array(tasks, dtype=api.ALIGN_TASK)
    bad, _, _ = check_align(g, reads, ref, tasks)
    assert not bad, bad[:10]
    g.close()


@pytest.mark.parametrize("band_mask,small,bandreg", [(None, None, None), ("0x0", "0", "0"), ("0xffff", "0", "0"), ("0xff", "0", "0xf")])
def test_gpu_similar_pairs_every_length(band_mask, small, bandreg, monkeypatch):
    """query = mutated target for every q from 1 to 560 (all register classes and their edges); with the
    default kernel mapping, with k_myers_small everywhere, with k_myers_band (plane store; banded + retry) everywhere,
    and with the plane store for q <= 128 plus k_myers_bandreg for every class above."""
    for k, v in (("LF_BAND_MASK", band_mask), ("LF_BANDREG_SMALL", small), ("LF_BANDREG", bandreg)):
        if v is not None:
            monkeypatch.setenv(k, v)
    rng = np.random.default_rng(77)
    ref = sim.make_reference(400_000, 9)
    reads, tasks, pos = [], [], 1000
    for L in range(1, 561):
        t = ref[pos:pos + L]
        q = sim.mutate_pair(t, float(rng.choice([0.05, 0.15, 0.25])), rng)
        reads.append(q)
        tasks.append((len(reads) - 1, 0, len(q), pos, L, 0, 0, 0))
        tasks.append((len(reads) - 1, 0, len(q), pos, L + 20, int(rng.integers(0, 2)) * 2, 1, 0))
        pos += L + 30
    tasks = np.array(tasks, dtype=api.ALIGN_TASK)
    g = api.LfGpu(sim.pack_pac(ref), len(ref))
    bad, _, _ = check_align(g, reads, ref, tasks)
    assert not bad, bad[:10]
    g.close()


def test_gpu_bad_tasks_are_flagged_and_empty_batch():
    ref = sim.make_reference(2000, 1)
    g = api.LfGpu(sim.pack_pac(ref), len(ref))
    read = ref[100:200].copy()
    tasks = np.array([(0, 0, 50, 0, 60, 0, 0, 0), (0, 90, 50, 0, 60, 0, 0, 0), (0, 0, 0, 0, 60, 0, 0, 0), (0, 0, 10, 1990, 60, 0, 0, 0), (3, 0, 10, 0, 60, 0, 0, 0)],
                     dtype=api.ALIGN_TASK)
    offs = np.array([0, 100], dtype=np.uint64)
    res, ops = g.align_batch(read, offs, tasks)
    assert int(res[0]["status"]) == 0 and [int(s) for s in res["status"][1:]] == [-3, -3, -3, -3]
    res, ops = g.align_batch(read, offs, tasks[:0])
    assert len(res) == 0
    g.close()


def test_gpu_config1_chunk_vs_oracle():
    """BASELINE config 1 shape (1 Mbp reference, 10 kbp reads at 15 %, SV mix): every round-1 task."""
    w = sim.make_workload(1_000_000, 60, 10_000, 0.15, 0.15, seed=1, sv_frac=0.3)
    tasks, chain, kind = workload_tasks(w)
    g = api.LfGpu(w.pac, len(w.ref))
    reads = [w.reads[w.read_off[i]:w.read_off[i + 1]] for i in range(w.n_reads)]
    bad, _, _ = check_align(g, reads, w.ref, tasks)
    assert not bad, bad[:10]
    g.close()


def test_gpu_extend_golden_and_random():
    cases = load_ksw()
    tcat, reads, tasks, off = [], [], [], 0
    for k, c in enumerate(cases):
        t = np.frombuffer(c["t"].encode(), dtype=np.uint8)
        tcat.append(t); reads.append(np.frombuffer(c["q"].encode(), dtype=np.uint8))
        p = c["prm"]
        tasks.append((k, 0, len(c["q"]), off, len(t), 0, 0, 0, p[0], p[1], p[2], p[3], p[4], p[5], len(c["q"])))
        off += len(t)
    ref = np.concatenate(tcat)
    g = api.LfGpu(sim.pack_pac(ref), len(ref))
    offs = np.zeros(len(reads) + 1, dtype=np.uint64); offs[1:] = np.cumsum([len(r) for r in reads])
    er = g.extend_batch(np.concatenate(reads), offs, np.array(tasks, dtype=api.EXTEND_TASK))
    for k, c in enumerate(cases):
        assert (int(er[k]["score"]), int(er[k]["qle"]), int(er[k]["tle"])) == (c["score"], c["qle"], c["tle"])
    # random, with strand / reversal flags, against the oracle
    rng = np.random.default_rng(8)
    w = sim.make_workload(300_000, 30, 4000, 0.15, 0.15, seed=6, sv_frac=0.5)
    et = []
    for k in range(400):
        rid = int(rng.integers(0, w.n_reads)); L = int(w.read_off[rid + 1] - w.read_off[rid])
        ql = int(rng.integers(20, min(L, 1800))); qo = int(rng.integers(0, L - ql + 1))
        to = int(rng.integers(0, len(w.ref) - ql - 40)) if k % 2 else int(w.chain(rid)[0][0])
        fl = int(rng.choice([0, 2])) | (1 if w.is_rev[rid] else 0)
        prm = (8, 1, 4, 1, 300, 200) if k % 7 == 0 else (0, 1, 0, 1, 40, 40) if k % 3 else (8, 1, 4, 1, 100, 200)  # w = 300: no shared-memory ring
        et.append((rid, qo, ql, to, ql + 20, fl, 0, 0) + prm + (ql,))
    et = np.array(et, dtype=api.EXTEND_TASK)
    g2 = api.LfGpu(w.pac, len(w.ref))
    er = g2.extend_batch(w.reads, w.read_off.astype(np.uint64), et)
    for i, t in enumerate(et):
        o = w.oriented(int(t["read_id"])) if int(t["flags"]) & 1 else w.reads[w.read_off[t["read_id"]]:w.read_off[t["read_id"] + 1]]
        q = o[t["q_off"]:t["q_off"] + t["q_len"]]; tt = w.ref[t["t_off"]:t["t_off"] + t["t_len"]]
        if int(t["flags"]) & 2:
            q, tt = q[::-1], tt[::-1]
        exp = O.oracle_extend(CODE[q].tobytes(), CODE[tt].tobytes(), int(t["o_del"]), int(t["e_del"]), int(t["o_ins"]), int(t["e_ins"]), int(t["w"]), int(t["zdrop"]))
        assert exp == (int(er[i]["score"]), int(er[i]["qle"]), int(er[i]["tle"])), (i, t)
    g.close(); g2.close()


def _validate_paths(w, tasks, res, ops, sample):
    """Size-independent properties: the op string is a valid alignment of the two slices whose cost
    equals the reported distance (so the distance is an upper bound reached by a real path)."""
    for i in sample:
        t, r = tasks[i], res[i]
        o = api.decode_ops(ops, int(r["ops_off"]), int(r["ops_len"]))
        nq = int(np.count_nonzero(o != 2)); nt = int(np.count_nonzero(o != 1))
        assert nq == int(t["q_len"]) and nt == int(r["end_location"]) + 1
        assert int(np.count_nonzero(o)) == int(r["edit_distance"])
        q, tt = task_strings(t, [w.reads[w.read_off[k]:w.read_off[k + 1]] for k in [int(t["read_id"])]] if False else _reads_list(w), w.ref)
        q = np.frombuffer(q, dtype=np.uint8); tt = np.frombuffer(tt, dtype=np.uint8)
        qi = np.cumsum(o != 2) - 1; ti = np.cumsum(o != 1) - 1
        diag = (o == 0) | (o == 3)
        eq = q[qi[diag]] == tt[ti[diag]]
        assert np.array_equal(eq, o[diag] == 0)


_RL = {}


def _reads_list(w):
    if id(w) not in _RL:
        _RL[id(w)] = [w.reads[w.read_off[k]:w.read_off[k + 1]] for k in range(w.n_reads)]
    return _RL[id(w)]


def test_gpu_full_size_properties():
    """BASELINE config 2 size (4.6 Mbp, 2k of its 20k x 10 kbp reads per run here): valid optimal-cost
    paths, exact agreement with the oracle on a sample, run-to-run identity, and independence from
    how the chunk is split into batches."""
    w = sim.make_workload(4_600_000, 2000, 10_000, 0.12, 0.15, seed=100, sv_frac=0.10)
    tasks, chain, kind = workload_tasks(w, long_ends_nopath=False)   # every task with its path, the junk heads / tails too
    g = api.LfGpu(w.pac, len(w.ref))
    ro = w.read_off.astype(np.uint64)
    res, ops = g.align_batch(w.reads, ro, tasks)
    assert int(np.count_nonzero(res["status"])) == 0
    rng = np.random.default_rng(1)
    sample = rng.choice(len(tasks), size=3000, replace=False)
    big = np.argsort(tasks["q_len"].astype(np.int64) * tasks["t_len"])[-40:]
    _validate_paths(w, tasks, res, ops, np.concatenate([sample, big]))
    reads = _reads_list(w)
    for i in np.concatenate([sample[:1500], big]):
        q, tt = task_strings(tasks[i], reads, w.ref)
        ed, end, path = O.oracle_align(q, tt, int(tasks[i]["mode"]))
        got = api.decode_ops(ops, int(res[i]["ops_off"]), int(res[i]["ops_len"])).tobytes()
        assert (ed, end, path) == (int(res[i]["edit_distance"]), int(res[i]["end_location"]), got), i
    # idempotence
    res2, ops2 = g.align_batch(w.reads, ro, tasks)
    assert res.tobytes() == res2.tobytes() and ops.tobytes() == ops2.tobytes()
    # batch-split independence: op strings and distances do not depend on batch composition
    half = len(tasks) // 2
    ra, oa = g.align_batch(w.reads, ro, tasks[:half])
    rb, ob = g.align_batch(w.reads, ro, tasks[half:])
    assert np.array_equal(np.concatenate([ra["edit_distance"], rb["edit_distance"]]), res["edit_distance"])
    assert np.array_equal(np.concatenate([ra["end_location"], rb["end_location"]]), res["end_location"])
    for i in sample[:500]:
        rr, oo, k = (ra, oa, i) if i < half else (rb, ob, i - half)
        assert np.array_equal(api.decode_ops(oo, int(rr[k]["ops_off"]), int(rr[k]["ops_len"])),
                              api.decode_ops(ops, int(res[i]["ops_off"]), int(res[i]["ops_len"])))
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("groupk,wide", [("7", "2048"), ("7", "0"), ("0", "2048")])
def test_gpu_group_classes(groupk, wide, monkeypatch):
    """k_myers_group (LANES lanes per task: path tasks of 257 .. 2048 rows, distance-only tasks up to 8192 rows) on both
    sides of every class boundary, NW / SHW mixed in a warp's bundle; LF_GROUPK=0 sends the same tasks to the older
    kernels.  Bit-exact against the oracle both ways, and the classes really ran."""
    from _common import group_class_batch
    monkeypatch.setenv("LF_GROUPK", groupk)
    monkeypatch.setenv("LF_GROUP_WIDE", wide)   # task count below which a class gives every task a whole warp
    ref, reads, tasks = group_class_batch()
    g = api.LfGpu(sim.pack_pac(ref), len(ref))
    bad, _, _ = check_align(g, reads, ref, tasks)
    assert not bad, bad[:10]
    cc = g.class_counts()
    if groupk == "7":
        assert all(cc.get(k, 0) > 0 for k in ("group_path16", "group_path32", "group_path64", "group_dist32", "group_dist64", "group_dist128", "group_dist256", "large")), cc
    else:
        assert not any(k.startswith("group") for k in cc), cc
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("bandreg", ["0xf", "0"])
def test_gpu_bandreg_near_diagonal(bandreg, monkeypatch):
    """k_myers_bandreg (sliding band in registers) on 4000 near-diagonal global tasks of 129..512 rows with every
    strand / direction flag, plus pairs whose path bulges out of the band (redone full width); and the same batch
    with the kernel switched off.  Bit-exact against the oracle both ways."""
    from test_emu_pipeline import _band_stress_batch, _near_diagonal_batch
    monkeypatch.setenv("LF_BANDREG", bandreg)
    rng = np.random.default_rng(31)
    ref = sim.make_reference(3_000_000, 8)
    reads, tasks = _near_diagonal_batch(rng, ref, 4000)
    r2, t2 = _band_stress_batch(rng, ref, 600)
    t2["read_id"] += len(reads)
    reads, tasks = reads + r2, np.concatenate([tasks, t2])
    g = api.LfGpu(sim.pack_pac(ref), len(ref))
    bad, _, _ = check_align(g, reads, ref, tasks)
    assert not bad, bad[:10]
    g.close()
