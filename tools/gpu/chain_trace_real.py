#!/usr/bin/env python
"""Phase trace of lf_gpu_align_chains on the config-2 chunk with the reference's chains (fixtures/config2.npz), and the
class timeline of the last alignment batch of the call (= round 3)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from lordfast_b200 import api, fixtures
fx = fixtures.load(sys.argv[1] if len(sys.argv) > 1 else "config2")
g = api.LfGpu(fx.pac, len(fx.ref))
pb = api.PinnedArray(g.lib, fx.reads.nbytes); hb = pb.view(np.uint8, len(fx.reads)); hb[:] = fx.reads
ro = fx.read_off.astype(np.uint64)
co, cl = np.array([0], dtype=np.int64), np.array([len(fx.ref)], dtype=np.int32)
if hasattr(fx, "contig_off") and fx.contig_off is not None:
    co, cl = fx.contig_off, fx.contig_len
for it in range(5):
    if it == 4:
        os.environ["LF_CHAIN_TRACE"] = os.environ.get("TRACE_LEVEL", "2")
    t0 = time.perf_counter()
    recs, text, st = g.align_chains(hb, ro, co, cl, fx.seeds, fx.chains, want_text=False)
    print("call %d: %.2f ms, %d records, round3 tasks %d" % (it, (time.perf_counter() - t0) * 1e3, len(recs), st.round3_tasks), flush=True)
st, en = g.class_timeline()
cc = g.class_counts()
print("last batch (round 3):")
for c in sorted(range(len(en)), key=lambda c: st[c]):
    if en[c] >= 0 and c != 17:
        print("  %-14s %6.3f -> %6.3f  tasks %d" % (api.CLASS_NAMES[c], st[c], en[c], cc.get(api.CLASS_NAMES[c], 0)))
