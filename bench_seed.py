#!/usr/bin/env python
"""Throughput of the FM-index seeding stage (lf_gpu_seed_batch) on one B200, next to the reference's own
getLocs_extend_whole_step (src/BWT.cpp:312-394) on the host cores.

    python bench_seed.py [--steps K] [--warmup W] [--reads N]

Workload: BASELINE configs[1] -- the 4.6 Mbp reference and the 20 000 x 10 kbp reads of fixtures/config2.npz (the same
reads bench.py aligns; a model workload of the same shape if the fixture is absent), lordFAST's own parameters
(MIN_ANCHOR_LEN 14, SAMPLING_COUNT 1000, MAX_REF_HITS 1000, k-mer table of 12).  One JSON line:
  value       read bases seeded per second, reads resident in HBM, seed lists delivered to pinned host memory
  e2e         the same through lf_gpu_seed_batch with the reads in (pinned) host memory
  kernels     CUDA-event times of search (positions, longest matches, filter, offsets) and locate (bwt_sa + strand split)
  roofline    the stage is random 64-byte reads of the bwt array (two per backward-search step, one per inverse-psi step):
              achieved = blocks read per second x 64 B against the HBM peak (MEASURED_PEAKS.json or the guide's fallback)
  cpu_baseline  the reference's code (oracle/_ref/libref_shim.so) on a sample of the reads, all host threads
The seed lists of the first call are checked against the reference for the CPU sample (the shim travels with the repo;
without it the check and the baseline are skipped and said so)."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from lordfast_b200 import api, fixtures, fmindex, sim  # noqa: E402

CODE = np.zeros(256, np.uint8)
for _i, _c in enumerate(b"ACGT"):
    CODE[_c] = _i


def workload(n_reads):
    path = os.path.join(ROOT, "fixtures", "config2.npz")
    if os.path.exists(path):
        fx = fixtures.load("config2")
        ref, reads, off, name = fx.ref, fx.reads, fx.read_off.astype(np.uint64), "config2_real reads (fixtures/config2.npz)"
    else:
        ref = sim.make_reference_dups(4_600_000, seed=2, dups=6)
        rng = np.random.default_rng(7)
        rl = []
        for i in range(20_000):
            a = int(rng.integers(0, len(ref) - 10_000))
            r = ref[a:a + 10_000].copy()
            hit = rng.random(10_000) < 0.13
            r[hit] = sim.ACGT[rng.integers(0, 4, size=int(hit.sum()))]
            rl.append(sim.revcomp(r) if i & 1 else r)
        reads, off, name = np.concatenate(rl), (np.arange(20_001) * 10_000).astype(np.uint64), "config2 model reads (substitutions only)"
    if n_reads and n_reads < len(off) - 1:
        off = off[:n_reads + 1]
        reads = reads[:int(off[-1])]
    return ref, np.ascontiguousarray(reads, dtype=np.uint8), off, name


def fm_for(ref):
    cache = os.path.join(ROOT, "fixtures", "config2_fm.npz")
    if os.path.exists(cache):
        z = np.load(cache)
        if int(z["l_pac"]) == len(ref):
            return fmindex.FmIndex(bwt=z["bwt"], sa=z["sa"], primary=int(z["primary"]), L2=z["L2"], seq_len=int(z["seq_len"]), sa_intv=int(z["sa_intv"]), l_pac=int(z["l_pac"]), k_cache=12)
    fm = fmindex.build(CODE[ref], k_cache=12)
    try:
        os.makedirs(os.path.dirname(cache), exist_ok=True)
        np.savez(cache, bwt=fm.bwt, sa=fm.sa, primary=fm.primary, L2=fm.L2, seq_len=fm.seq_len, sa_intv=fm.sa_intv, l_pac=fm.l_pac)
    except OSError:
        pass
    return fm


def hbm_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_GBps", "hbm_copy_gbs", "hbm_bw_gbs"):
                if k in d:
                    return float(d[k]), "MEASURED_PEAKS.json:" + k
            for k, v in d.items():
                if "hbm" in k.lower() and isinstance(v, (int, float)):
                    return float(v), "MEASURED_PEAKS.json:" + k
        except Exception:
            pass
    return 6550.0, "B200_PROFILING.md fallback (measured copy bandwidth 6.55 TB/s)"


def cpu_reference(ref, reads, off, sample, gpu_lists):
    """the reference's getLocs_extend_whole_step on `sample` reads, all host threads; also checks the GPU lists of those reads"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import _oracle
        if not (_oracle.have_ref() and hasattr(_oracle.ref(), "ref_fm_load")):
            return None, "oracle/_ref/libref_shim.so (with the seeding exports) is not on this box"
    except Exception as e:   # noqa: BLE001
        return None, "reference shim unavailable: %s" % e
    import ctypes as C
    lib = _oracle.ref()
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "r.fa")
        _oracle.write_fasta(fa, ref)
        _oracle.ref_fm_load(fa, 12)
    rb = reads.tobytes()
    n = min(sample, len(off) - 1)
    qs = [rb[int(off[i]):int(off[i + 1])] + b"\0" for i in range(n)]
    arr = (C.c_char_p * n)(*qs)
    ql = (C.c_uint32 * n)(*[len(q) - 1 for q in qs])
    threads = os.cpu_count() or 1
    hits = C.c_uint64()
    lib.ref_fm_seed_batch(min(n, 2 * threads), arr, ql, 1000, 14, 1000, threads, C.byref(hits))   # warm-up
    sec = lib.ref_fm_seed_batch(n, arr, ql, 1000, 14, 1000, threads, C.byref(hits))
    bases = sum(len(q) - 1 for q in qs)
    fwd, fo, rev, ro = gpu_lists
    mism = 0
    for i in range(min(n, 64)):
        f, r = _oracle.ref_fm_seed(qs[i][:-1])
        mism += int(not np.array_equal(f, fwd[int(fo[i]):int(fo[i + 1])])) + int(not np.array_equal(r, rev[int(ro[i]):int(ro[i + 1])]))
    return {"value": bases / sec / 1e6, "unit": "Mbp/s", "cores": threads, "kind": "reference",
            "sample": "%d reads / %.1f Mbp through the reference's getLocs_extend_whole_step (oracle/_ref), %d threads; %d seeds" % (n, bases / 1e6, threads, hits.value),
            "gpu_lists_checked": min(n, 64), "gpu_list_mismatches": mism}, None


def measure(steps=10, warmup=3, n_reads=0, cpu_sample=400, cpu_baseline=True, data=None):
    import torch   # device memory / stream plumbing of the process; the kernels are the library's
    if not torch.cuda.is_available():
        raise SystemExit("bench_seed.py needs a CUDA device (there is no CPU path for the seeding stage)")
    ref, reads, off, name = data if data is not None else workload(n_reads)   # data: (reference ASCII, read bases, offsets, name) already in memory
    fm = fm_for(ref)
    g = api.LfGpu(sim.pack_pac(ref), len(ref))
    t0 = time.perf_counter()
    g.seed_init(fm)
    init_ms = (time.perf_counter() - t0) * 1e3
    pin = api.PinnedArray(g.lib, reads.nbytes)
    hb = pin.view(np.uint8, len(reads)); hb[:] = reads
    bases = int(off[-1])
    first = g.seed_batch(hb, off)
    for _ in range(max(warmup, 3) - 1):
        g.seed_batch(None, off, copy=False)
    t = []
    for _ in range(steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.seed_batch(None, off, copy=False)
        t.append(time.perf_counter() - t0)
    st = g.seed_stats()
    te = []
    for _ in range(steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.seed_batch(hb, off, copy=False)
        te.append(time.perf_counter() - t0)
    ms, ms_e = 1e3 * float(np.median(t)), 1e3 * float(np.median(te))
    blocks = 2 * st["search_steps"] + st["locate_steps"]
    kern_ms = st["search_ms"] + st["locate_ms"]
    peak, src = hbm_peak_gbs()
    ach = blocks * 64 / (kern_ms * 1e-3) / 1e9
    line = {"metric": "seeded Mbp/s (FM-index seeding stage, getLocs_extend_whole_step)", "value": bases / (ms * 1e-3) / 1e6, "unit": "Mbp/s", "n_gpus": 1,
            "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms, "higher_is_better": True, "dtype": "u64", "data": "synthetic",
            "config": {"workload": name, "reads": len(off) - 1, "read_bases": bases, "reference_bp": len(ref), "min_anchor_len": 14, "sampling_count": 1000, "max_ref_hits": 1000,
                       "k_cache": 12, "sa_intv": fm.sa_intv, "l2": "index (bwt %.1f MB, sa %.1f MB, k-mer table 268 MB) + 202 MB of reads: the k-mer table and the reads exceed L2, the bwt of a 4.6 Mbp reference does not"
                       % (fm.bwt.nbytes / 1e6, fm.sa.nbytes / 1e6)},
            "e2e": {"value": bases / (ms_e * 1e-3) / 1e6, "unit": "Mbp/s", "ms_per_step": ms_e, "h2d_bytes_per_step": int(reads.nbytes + off.nbytes),
                    "d2h_bytes_per_step": int(12 * (len(first[0]) + len(first[2])) + 2 * off.nbytes), "call": "lf_gpu_seed_batch"},
            "kernels": {"search_ms": st["search_ms"], "locate_ms": st["locate_ms"], "positions": st["positions"], "hits": st["hits"],
                        "search_steps": st["search_steps"], "locate_steps": st["locate_steps"], "index_upload_and_table_ms": init_ms},
            "roofline": {"bound": "hbm (random 64-byte blocks)", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": src,
                         "blocks_per_step": blocks, "note": "2 bwt blocks per backward-search step + 1 per inverse-psi step, counted by the kernels; a 4.6 Mbp index is L2-resident, so this is an L2/latency figure, not DRAM"},
            "gpu_launches": 8}
    if cpu_baseline:
        cb, why = cpu_reference(ref, reads, off, cpu_sample, first)
        line["cpu_baseline"] = cb if cb else {"unavailable": why}
    pin.free()
    g.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reads", type=int, default=0, help="first N reads of the workload (0 = all 20 000)")
    ap.add_argument("--cpu-sample", type=int, default=400)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    print(json.dumps(measure(a.steps, a.warmup, a.reads, a.cpu_sample, not a.no_cpu_baseline)))


if __name__ == "__main__":
    main()
