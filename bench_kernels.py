#!/usr/bin/env python
"""Kernel microbenchmark (BASELINE.json configs[4]): batched gap alignment sweep, target length
L in {100 .. 10 000} bp, query = target mutated at divergence d in {5, 10, 15, 20} %, NW with path and an
SHW variant with 20 extra target bases; plus ksw_extend2 end-extensions (identical prefix of L/2, then
random; lordFAST's clip and split parameter sets).  Reports GCUPS (full-matrix cells q*t per second of
kernel time, CUDA events on the library's stream) and, for the alignments, the fraction of the INT32
roofline at 16 ops per (32-row word x column).  Inputs are resident in HBM; a sample of every point is
checked against the oracle.

    python bench_kernels.py [--pairs 65536] [--quick]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=65536)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--check", type=int, default=24, help="pairs per point verified against the oracle")
    ap.add_argument("--lengths", type=int, nargs="*", help="target lengths instead of the config-5 list")
    ap.add_argument("--divs", type=float, nargs="*", help="divergences instead of the config-5 list")
    ap.add_argument("--nw-only", action="store_true", help="skip the SHW variant and the ksw_extend2 points")
    a = ap.parse_args()
    import _oracle as O
    from lordfast_b200 import api, sim

    rng = np.random.default_rng(2024)
    ref_len = 8_000_000
    ref = sim.make_reference(ref_len, 7)
    g = api.LfGpu(sim.pack_pac(ref), ref_len)
    peak = g.int32_peak(1)
    lengths = [100, 200, 500, 1000, 2000, 5000, 10000]
    divs = [0.05, 0.10, 0.15, 0.20]
    if a.quick:
        lengths, divs = [100, 500, 2000], [0.15]
    if a.lengths:
        lengths = a.lengths
    if a.divs:
        divs = a.divs
    rows = []
    for L in lengths:
        n = a.pairs   # 2^16 pairs at every length
        n_src = min(n, 4096)  # distinct pairs; tasks cycle over them (inputs stay far larger than L2 for small L)
        for d in divs:
            starts = rng.integers(0, ref_len - L - 64, size=n_src)
            reads = [sim.mutate_pair(ref[s:s + L], d, rng) for s in starts]
            off = np.zeros(n_src + 1, dtype=np.uint64); off[1:] = np.cumsum([len(r) for r in reads])
            bases = np.concatenate(reads)
            qlen = np.diff(off).astype(np.uint32)
            for mode, extra in (((api.LF_MODE_NW, 0),) if a.nw_only else ((api.LF_MODE_NW, 0), (api.LF_MODE_SHW, 20))):
                t = np.zeros(n, dtype=api.ALIGN_TASK)
                idx = np.arange(n) % n_src
                t["read_id"], t["q_off"], t["q_len"] = idx, 0, qlen[idx]
                t["t_off"], t["t_len"], t["mode"] = starts[idx], L + extra, mode
                g.upload_reads(bases, off)
                g.upload_align_tasks(t)
                g.run_align(); g.sync()
                ms = []
                for _ in range(3 if L < 5000 else 1):
                    g.run_align(); g.sync()
                    ms.append(g.stats().last_main_kernel_ms)
                ms = float(np.min(ms))
                cells = float(t["q_len"].astype(np.int64) @ t["t_len"].astype(np.int64))
                wc = float((((t["q_len"].astype(np.int64) + 31) // 32) * t["t_len"].astype(np.int64)).sum())
                res = np.zeros(n, dtype=api.ALIGN_RESULT)
                ops = np.zeros(g.lib.lf_gpu_ops_capacity(t.ctypes.data, n), dtype=np.uint8)
                g.download_align(res, ops)
                bad = 0
                for k in range(min(a.check, n_src)):
                    q = reads[k].tobytes(); tt = ref[starts[k]:starts[k] + L + extra].tobytes()
                    ed, end, path = O.oracle_align(q, tt, mode)
                    got = api.decode_ops(ops, int(res[k]["ops_off"]), int(res[k]["ops_len"])).tobytes()
                    bad += (ed, end, path) != (int(res[k]["edit_distance"]), int(res[k]["end_location"]), got)
                row = {"kind": "NW" if mode == 0 else "SHW", "L": L, "div": d, "pairs": n, "kernel_ms": round(ms, 3),
                       "gcups": round(cells / ms / 1e6, 1), "int32_tops": round(16 * wc / ms / 1e9, 3),
                       "roofline_frac": round(16 * wc / ms / 1e9 / peak, 3), "oracle_mismatches": bad}
                rows.append(row)
                print(json.dumps(row), flush=True)
    # ksw_extend2
    code = np.zeros(256, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        code[ch] = i
    for L in ([] if a.nw_only else [500, 2000] if a.quick else [500, 2000, 10000]):
        n = min(a.pairs, 16384 if L >= 10000 else a.pairs)
        n_src = min(n, 2048)
        starts = rng.integers(0, ref_len - L - 64, size=n_src)
        reads = [np.concatenate([ref[s:s + L // 2], sim.ACGT[rng.integers(0, 4, size=L - L // 2, dtype=np.uint8)]]) for s in starts]
        off = np.zeros(n_src + 1, dtype=np.uint64); off[1:] = np.cumsum([len(r) for r in reads])
        bases = np.concatenate(reads)
        for name, prm in (("clip", (0, 1, 0, 1, 40, 40)), ("split", (8, 1, 4, 1, 100, 200))):
            et = np.zeros(n, dtype=api.EXTEND_TASK)
            idx = np.arange(n) % n_src
            et["read_id"], et["q_off"], et["q_len"], et["t_off"], et["t_len"] = idx, 0, L, starts[idx], L
            et["o_del"], et["e_del"], et["o_ins"], et["e_ins"], et["w"], et["zdrop"], et["h0"] = (*prm, L)
            import time
            g.upload_reads(bases, off)
            er = g.extend_batch(bases, off, et)
            t0 = time.perf_counter()
            for _ in range(2):
                er = g.extend_batch(bases, off, et)
            wall = (time.perf_counter() - t0) / 2 * 1e3
            w = prm[4]
            band_cells = float(np.minimum(er["tle"].astype(np.int64) + prm[5], L).sum()) * (2 * w + 1)
            bad = 0
            for k in range(min(a.check, n_src)):
                exp = O.oracle_extend(code[reads[k]].tobytes(), code[ref[starts[k]:starts[k] + L]].tobytes(), *prm)
                bad += exp != (int(er[k]["score"]), int(er[k]["qle"]), int(er[k]["tle"]))
            row = {"kind": "ksw_extend2_" + name, "L": L, "pairs": n, "wall_ms_incl_copies": round(wall, 2),
                   "gcups_band_cells": round(band_cells / wall / 1e6, 2), "oracle_mismatches": bad}
            rows.append(row)
            print(json.dumps(row), flush=True)
    print(json.dumps({"int32_peak_tops": peak, "rows": len(rows), "mismatches": int(sum(r["oracle_mismatches"] for r in rows))}))
    g.close()


if __name__ == "__main__":
    main()
