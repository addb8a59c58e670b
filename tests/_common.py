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
