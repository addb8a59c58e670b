"""Helpers shared by the CPU (emulator) and GPU parity tests."""
import json
import os
import subprocess

import numpy as np

import _oracle as O
from lordfast_b200 import api, sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "liblfgpu_emu_testonly.so")
CODE = np.zeros(256, dtype=np.uint8)
CODE[:] = 4
for _i, _c in enumerate(b"ACGT"):
    CODE[_c] = _i
    CODE[_c + 32] = _i


def build_emu():
    """g++ build of the pipeline on the fiber emulator (test-only debugging aid, not a product path)."""
    srcs = [os.path.join(EMU_DIR, "lfgpu_emu.cpp"), os.path.join(EMU_DIR, "cuda_emu.cpp")]
    deps = srcs + [os.path.join(EMU_DIR, "cuda_emu.h")] + [os.path.join(ROOT, "lordfast_b200", "csrc", f) for f in
                                                          ("lf_kernels.cuh", "lf_pipeline.inl", "lf_backend.h", "lf_chain.inl")] + [os.path.join(ROOT, "include", "lf_gpu.h")]
    deps = [d for d in deps if os.path.exists(d)]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-I" + os.path.join(ROOT, "include"),
                               "-I" + EMU_DIR, "-o", EMU_LIB] + srcs)
    return EMU_LIB


def load_pairs():
    return json.load(open(os.path.join(GOLDEN, "align_pairs.json")))


def load_ksw():
    return json.load(open(os.path.join(GOLDEN, "ksw_cases.json")))


def load_chains():
    z = np.load(os.path.join(GOLDEN, "chains.npz"))
    return z, json.load(open(os.path.join(GOLDEN, "chains.json")))


def task_strings(t, reads_list, ref):
    """The two byte strings edlibAlign would be handed for this task."""
    r = reads_list[int(t["read_id"])]
    o = sim.revcomp(r) if int(t["flags"]) & api.LF_F_READ_REV else r
    q = o[int(t["q_off"]):int(t["q_off"]) + int(t["q_len"])]
    tt = ref[int(t["t_off"]):int(t["t_off"]) + int(t["t_len"])]
    if int(t["flags"]) & api.LF_F_REVERSE_BOTH:
        q, tt = q[::-1], tt[::-1]
    if int(t["flags"]) & api.LF_F_RC_QUERY:
        q = sim.revcomp(q)
    return q.tobytes(), tt.tobytes()


def check_align(g, reads_list, ref, tasks):
    """Runs tasks through lf_gpu_align_batch and compares every field with the oracle."""
    offs = np.zeros(len(reads_list) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads_list])
    res, ops = g.align_batch(np.concatenate(reads_list), offs, tasks)
    bad = []
    for i, t in enumerate(tasks):
        q, tt = task_strings(t, reads_list, ref)
        want = not (int(t["flags"]) & api.LF_F_NO_PATH)
        ed, end, path = O.oracle_align(q, tt, int(t["mode"]), want)
        got = api.decode_ops(ops, int(res[i]["ops_off"]), int(res[i]["ops_len"])).tobytes() if want else b""
        if (int(res[i]["edit_distance"]), int(res[i]["end_location"]), got, int(res[i]["status"])) != (ed, end, path if want else b"", 0):
            bad.append(i)
    return bad, res, ops


def pairs_as_batch(pairs):
    """Golden (q, t) strings -> a reference made of the targets, one read per query, one task per pair."""
    tcat, reads, tasks, off = [], [], [], 0
    for k, p in enumerate(pairs):
        t = np.frombuffer(p["t"].encode(), dtype=np.uint8)
        # golden targets are ACGT only, so they can live in a 2-bit reference
        tcat.append(t)
        reads.append(np.frombuffer(p["q"].encode(), dtype=np.uint8))
        tasks.append((k, 0, len(p["q"]), off, len(t), 0, p["mode"], 0))
        off += len(t)
    ref = np.concatenate(tcat)
    return ref, reads, np.array(tasks, dtype=api.ALIGN_TASK)


def craft_chains(ref, seed=3):
    """Reads stitched from reference pieces with 400 bp deletions between them, with sparse exact anchors.
    Pieces holding a single anchor make records of fewer than two anchors, which the reference drops
    (src/LordFAST.cpp:1991 / :2063); error-free stretches make anchor + gap units that are matches only, so
    CIGAR / MD runs merge across many anchors.  Returns [(read, [(tPos, qPos, len), ...]), ...]."""
    rng = np.random.default_rng(seed)
    cases = []
    layouts = [[1500, 40, 1500], [40, 1500], [1500, 40], [1200, 130, 1200], [40, 40, 1500], [900, 40, 40, 900], [2000], [700, 700]]
    for li, lay in enumerate(layouts):
        for noisy in (0, 1, 2):
            tpos = int(rng.integers(1000, 3000)) + 6000 * (li % 4)
            read, seeds, qpos = [], [], 0
            for L in lay:
                piece = ref[tpos:tpos + L].copy()
                anchors = [(L - 14) // 2] if L <= 60 else [10, L - 30] if L <= 140 else list(range(5, L - 20, 55))
                prot = np.zeros(L, bool)
                for a in anchors:
                    prot[a:a + 14] = True
                if noisy:
                    cand = np.flatnonzero(~prot)
                    k = max(1, len(cand) // (12 if noisy == 1 else 40))
                    for x in rng.choice(cand, size=min(k, len(cand)), replace=False):
                        piece[x] = sim.ACGT[(int(np.searchsorted(sim.ACGT, piece[x])) + 1 + int(rng.integers(0, 3))) % 4]
                seeds += [(tpos + a, qpos + a, 14) for a in anchors]
                read.append(piece)
                qpos += L
                tpos += L + 400
            q = np.concatenate(read)
            if noisy == 2:  # junk at both ends: the head / tail alignments and their clips take part
                h, t = sim.ACGT[rng.integers(0, 4, size=30, dtype=np.uint8)], sim.ACGT[rng.integers(0, 4, size=25, dtype=np.uint8)]
                q = np.concatenate([h, q, t])
                seeds = [(a, b + 30, c) for a, b, c in seeds]
            cases.append((q, seeds))
    return cases


CRAFTED_REF = (40_000, 12)  # sim.make_reference(length, seed) of the crafted chains


def load_crafted():
    return json.load(open(os.path.join(GOLDEN, "chains_crafted.json")))


def group_class_batch(seed=21, big=True):
    """Tasks for every k_myers_group class (lordfast_b200/csrc/lf_kernels.cuh): path tasks of 257 .. 2048 rows off the
    diagonal, in prefix mode and near the diagonal, distance-only tasks of 513 .. 8192+ rows, lengths on both sides of
    every class boundary, NW and SHW mixed inside a warp's bundle, an SHW task whose best prefix is the empty one."""
    rng = np.random.default_rng(seed)
    ref = sim.make_reference(60000, 8)
    reads, tasks = [], []
    junk = lambda m: sim.ACGT[rng.integers(0, 4, size=m, dtype=np.uint8)]
    qlens = [257, 300, 384, 385, 512, 513, 700, 1000, 1024, 1025, 1500, 2047, 2048]
    for k, ql in enumerate(qlens):
        s = int(rng.integers(100, 30000))
        similar = sim.mutate_pair(ref[s:s + int(ql / 1.05)], 0.15, rng)[:ql]
        if len(similar) < ql:
            similar = np.concatenate([similar, junk(ql - len(similar))])
        for q, to, tl, flags, mode in (
                (similar, s, int(ql / 1.05), 0, 0),                           # near the diagonal
                (similar, s, int(ql / 1.05) + 20, 0, 1),                      # prefix mode, similar
                (junk(ql), int(rng.integers(0, 30000)), ql + 20, 2, 1),       # junk head: prefix mode, both reversed
                (junk(ql), int(rng.integers(0, 30000)), max(2, ql // 3), 1, 0),   # off the diagonal, reverse strand
                (junk(ql), int(rng.integers(0, 30000)), ql + 20, 8, 1),       # distance only
                (similar, s, int(ql / 1.05), 8 | 4, 0)):                      # distance only, global, RC query
            if not O.is_leaf(len(q), tl) and not (flags & 8):
                continue
            reads.append(q)
            tasks.append((len(reads) - 1, 0, len(q), to, tl, flags, mode, 0))
    if big:
        for ql, tl, flags, mode in ((2049, 2069, 8, 1), (4096, 300, 8, 0), (4097, 200, 8 | 2, 1), (8192, 150, 8, 1), (8193, 150, 8, 1), (3000, 100, 0, 0), (5000, 60, 0, 1)):
            reads.append(junk(ql))
            tasks.append((len(reads) - 1, 0, ql, int(rng.integers(0, 30000)), tl, flags, mode, 0))
    # the empty prefix is the best one: an all-T query against an all-A stretch cannot exist in a random reference, so plant it
    ref[50000:50400] = ord("A")
    reads.append(np.full(600, ord("T"), dtype=np.uint8))
    tasks.append((len(reads) - 1, 0, 600, 50000, 400, 0, 1, 0))
    return ref, reads, np.array(tasks, dtype=api.ALIGN_TASK)
