"""Experiment: how much does lf_gpu_align_chains gain when a chunk is cut into L parts that run concurrently on
L contexts of the same GPU (host phases of one part overlapping GPU / PCIe phases of another)?"""
import ctypes as C
import sys
import threading
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from lordfast_b200 import api, sim  # noqa: E402

ref_len, n_reads, read_len, e0, e1 = bench.WORKLOADS["config2_4.6Mbp_20kx10k"]
w = sim.make_workload(ref_len, n_reads, read_len, e0, e1, seed=1, sv_frac=0.1)
seeds_a, chains_a = api.workload_chains(w)
read_off = w.read_off.astype(np.uint64)
total_bases = int(read_off[-1])
cg_off, cg_len = w.contig_off, w.contig_len


def parts(L):
    out = []
    n = len(chains_a)
    for l in range(L):
        lo, hi = n * l // L, n * (l + 1) // L
        ch = chains_a[lo:hi].copy()
        rmin, rmax = int(ch["read_id"].min()), int(ch["read_id"].max())
        smin = int(ch["seed_off"].min()); smax = int((ch["seed_off"] + ch["n_seeds"]).max())
        ch["read_id"] -= rmin; ch["seed_off"] -= smin
        offs = (read_off[rmin:rmax + 2] - read_off[rmin]).astype(np.uint64)
        out.append(dict(chains=ch, seeds=np.ascontiguousarray(seeds_a[smin:smax]), b0=int(read_off[rmin]), b1=int(read_off[rmax + 1]), offs=offs))
    return out


for L in (1, 2, 3, 4):
    ps = parts(L)
    gs = [api.LfGpu(w.pac, len(w.ref)) for _ in range(L)]
    # reads in pinned host memory, like bench.py
    for p, g in zip(ps, gs):
        pa = api.PinnedArray(g.lib, p["b1"] - p["b0"])
        hb = pa.view(np.uint8); p["_keep"] = pa
        hb[:] = w.reads[p["b0"]:p["b1"]]
        p["bases"] = hb
        p["reads"] = api.Reads(hb.ctypes.data, p["offs"].ctypes.data, len(p["offs"]) - 1)
        p["cg"] = api.Contigs(cg_off.ctypes.data, cg_len.ctypes.data, len(cg_off))

    def call(p, g):
        out = C.c_void_p()
        rc = g.lib.lf_gpu_align_chains(g.ctx, C.byref(p["reads"]), C.byref(p["cg"]), p["seeds"].ctypes.data, p["chains"].ctypes.data, len(p["chains"]),
                                       g.pac.ctypes.data, C.byref(out))
        assert rc == 0, rc
        g.lib.lf_chain_results_free(out)

    def step():
        th = [threading.Thread(target=call, args=(p, g)) for p, g in zip(ps[1:], gs[1:])]
        for t in th: t.start()
        call(ps[0], gs[0])
        for t in th: t.join()
    for _ in range(3): step()
    t0 = time.perf_counter()
    K = 8
    for _ in range(K): step()
    ms = (time.perf_counter() - t0) / K * 1e3
    print(f"lanes {L}: {ms:.2f} ms per chunk  -> {total_bases / ms / 1e3:.0f} Mbp/s", flush=True)
    for g in gs: g.close()
