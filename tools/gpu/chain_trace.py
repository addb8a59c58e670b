#!/usr/bin/env python
"""Phase trace of lf_gpu_align_chains on a config-2 chunk (LF_CHAIN_TRACE=1 prints the host-side phase times)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from lordfast_b200 import api, sim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
w = sim.make_workload(4_600_000, n, 10_000, 0.12, 0.15, seed=100, sv_frac=0.10)
g = api.LfGpu(w.pac, len(w.ref))
seeds, chains = api.workload_chains(w)
pb = api.PinnedArray(g.lib, w.reads.nbytes); hb = pb.view(np.uint8, len(w.reads)); hb[:] = w.reads
ro = w.read_off.astype(np.uint64)
for it in range(5):
    if it == 4:
        os.environ["LF_CHAIN_TRACE"] = os.environ.get("TRACE_LEVEL", "1")
    t0 = time.perf_counter()
    recs, text, st = g.align_chains(hb, ro, w.contig_off, w.contig_len, seeds, chains, want_text=False)
    print("call %d: %.2f ms, %d records" % (it, (time.perf_counter() - t0) * 1e3, len(recs)), flush=True)
