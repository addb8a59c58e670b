#!/usr/bin/env python
"""Steady-state throughput of lf_gpu_align_chains with P contexts in flight (each call = one whole config-2 chunk,
host buffers in, records + text out): P host threads, each with its own context, calling back to back."""
import ctypes as C, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from lordfast_b200 import api, sim
n = int(os.environ.get("NREADS", "20000"))
w = sim.make_workload(4_600_000, n, 10_000, 0.12, 0.15, seed=100, sv_frac=0.10)
seeds, chains = api.workload_chains(w)
g0 = api.LfGpu(w.pac, len(w.ref))
pb = api.PinnedArray(g0.lib, w.reads.nbytes); hb = pb.view(np.uint8, len(w.reads)); hb[:] = w.reads
ro = w.read_off.astype(np.uint64)
rs = api.Reads(hb.ctypes.data, ro.ctypes.data, w.n_reads)
cg = api.Contigs(w.contig_off.ctypes.data, w.contig_len.ctypes.data, 1)
def call(g):
    out = C.c_void_p()
    rc = g.lib.lf_gpu_align_chains(g.ctx, C.byref(rs), C.byref(cg), seeds.ctypes.data, chains.ctypes.data, len(chains), g.pac.ctypes.data, C.byref(out))
    assert rc == 0, rc
    nr = C.c_size_t(); g.lib.lf_chain_results_records(out, C.byref(nr)); g.lib.lf_chain_results_free(out)
    return nr.value
for P in [int(x) for x in sys.argv[1:]] or [1, 2, 3, 4]:
    ctxs = [g0] + [api.LfGpu(w.pac, len(w.ref)) for _ in range(P - 1)]
    os.environ["LF_HOST_THREADS"] = str(max(2, (os.cpu_count() or 4) // P))
    per = 8
    def worker(g, k):
        for _ in range(k): call(g)
    for g in ctxs: call(g); call(g)
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(g, per)) for g in ctxs]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    print("P=%d lanes=%s: %.2f ms per chunk (%.1f Gbp/s), %.2f ms per call" % (P, os.environ.get("LF_CHAIN_LANES", "auto"), dt / (P * per) * 1e3, w.total_bases / (dt / (P * per)) / 1e9, dt / per * 1e3), flush=True)
    for g in ctxs[1:]: g.close()
