#!/usr/bin/env python
"""bench.py -- aligned Mbp/s of the per-candidate alignment stage on B200 (BASELINE.json metric).

A step = one pass of the stage over one chunk of synthetic reads (configs[1]: 4.6 Mbp reference,
20k x 10 kbp reads at 12-15 % error, one candidate chain per read, 10 % SV mix) whose round-1 tasks
(head SHW, gap NW, tail SHW, all with CIGAR path) are resident in HBM when the timed region starts.

  value      whole-job Mbp/s, kernels only (prep, sort, scans, all alignment kernels; inputs in HBM)
  e2e        the same through lf_gpu_align_batch with pinned HOST buffers: H2D of reads+tasks and
             D2H of results + 2-bit op stream inside the timed region
  roofline   k_myers_small (the dominant kernels): 16 INT32 ops per (32-row word x column),
             single-pass full-matrix count, against the LOP3/IADD3 issue rate
             measured live by lf_gpu_int32_peak (MEASURED_PEAKS.json has no INT32 figure)
  cpu_baseline  the reference's own alignChain_edlib (oracle/_ref) replayed over the same chains on
             the host cores (kind "reference"), or the oracle port if oracle/_ref is absent

`--impl reference` times only that CPU arm.  N > 1: one process per GPU (torchrun), reads sharded by
rank (weak scaling: every rank gets a chunk of the same size), no collective on the data path.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # one hardware queue per size-class stream; before torch starts CUDA

WORKLOADS = {
    # name: (ref_len, n_reads, read_len, err_lo, err_hi)
    "config1_1Mbp_200x10k": (1_000_000, 200, 10_000, 0.15, 0.15),
    "config2_4.6Mbp_20kx10k": (4_600_000, 20_000, 10_000, 0.12, 0.15),
    "config2_small_4.6Mbp_2kx10k": (4_600_000, 2_000, 10_000, 0.12, 0.15),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2_4.6Mbp_20kx10k", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sv-frac", type=float, default=0.10, help="fraction of SV/chimera reads (0.10 = the configured mix)")
    ap.add_argument("--e2e-calls", type=int, default=0, help="e2e: the chunk goes through this many concurrent lf_gpu_align_chains calls (contexts) per step; 1 = one call; "
                    "0 = one per 4 host cores this rank can count on, at most 4 (measured: 4 calls 14.4 vs 16.4 ms on 16 cores / 1 GPU, but 32.9 vs 29.3 ms on 32 cores / 8 GPUs)")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(name, seed, sv_frac=0.10):
    from lordfast_b200 import sim
    from lordfast_b200.chain_tasks import workload_tasks
    ref_len, n_reads, read_len, e0, e1 = WORKLOADS[name]
    w = sim.make_workload(ref_len, n_reads, read_len, e0, e1, seed=seed, sv_frac=sv_frac)
    tasks, chain, kind = workload_tasks(w)
    return w, tasks


def cpu_reference_rate(w, nthreads, max_chains=None):
    """Mbp/s of the reference's alignChain_edlib (oracle/_ref) over this workload's chains."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    n = w.n_reads if max_chains is None else min(w.n_reads, max_chains)
    if O.have_ref():
        lib = O.ref()
        idx_off = (C.c_int64 * 1)(0)
        idx_len = (C.c_int32 * 1)(len(w.ref))
        lib.ref_set_index(w.pac.ctypes.data, len(w.ref), 1, idx_off, idx_len)
        oriented = [np.ascontiguousarray(w.oriented(i)).tobytes() for i in range(n)]
        qptr = (C.c_char_p * n)(*oriented)
        seeds = np.ascontiguousarray(w.seeds[: int(w.seed_off[n])], dtype=np.uint32)
        seed_off = np.ascontiguousarray(w.seed_off[: n + 1], dtype=np.int64)
        rl = np.array([len(o) for o in oriented], dtype=np.int32)
        isrev = np.ascontiguousarray(w.is_rev[:n], dtype=np.uint8)
        nsam = C.c_int64()
        lib.ref_replay_chains.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        secs = lib.ref_replay_chains(n, seed_off.ctypes.data, seeds.ctypes.data, C.cast(qptr, C.c_void_p), rl.ctypes.data,
                                     isrev.ctypes.data, nthreads, C.byref(nsam))
        bases = int(rl.sum())
        return bases / secs / 1e6, "reference", nthreads, f"{n} chains / {bases / 1e6:.1f} Mbp through the reference's alignChain_edlib (oracle/_ref), {nthreads} threads"
    # oracle port, single thread, small sample
    n = min(n, 100)
    idx = O.RefIndex(w.ref.tobytes())
    t0 = time.time()
    bases = 0
    for i in range(n):
        seeds = [tuple(int(x) for x in s) for s in w.chain(i)]
        q = w.oriented(i).tobytes()
        O.oracle_align_chain(idx, seeds, q, int(w.is_rev[i]))
        bases += len(q)
    secs = time.time() - t0
    return bases / secs / 1e6, "port", 1, f"{n} chains / {bases / 1e6:.1f} Mbp through oracle/lf_oracle.c (scalar O(q*t) port), 1 thread"


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl_name = a.workload

    if a.impl == "reference":
        if rank != 0:
            return
        w, tasks = make_inputs("config2_small_4.6Mbp_2kx10k" if wl_name.startswith("config2") else wl_name, seed=100)
        nthreads = os.cpu_count() or 1
        rates = []
        for it in range(a.warmup + a.steps):
            r, kind, cores, sample = cpu_reference_rate(w, nthreads)
            if it >= a.warmup:
                rates.append(r)
        v = float(np.mean(rates))
        ms = w.total_bases / 1e6 / v * 1e3
        print(json.dumps({"impl": "reference", "metric": "aligned Mbp/s (alignment stage)", "value": v, "unit": "Mbp/s", "n_gpus": a.gpus,
                          "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                          "config": {"workload": wl_name, "sample": sample},
                          "cpu_baseline": {"value": v, "unit": "Mbp/s", "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": v, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the alignment stage has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from lordfast_b200 import api

    w, tasks = make_inputs(wl_name, seed=100 + rank, sv_frac=a.sv_frac)  # every rank gets its own chunk of the same size
    g = api.LfGpu(w.pac, len(w.ref))
    n = len(tasks)
    total_bases = w.total_bases
    read_off = w.read_off.astype(np.uint64)

    # pinned host staging for the e2e arm
    cap = g.lib.lf_gpu_ops_capacity(tasks.ctypes.data, n)
    p_tasks = api.PinnedArray(g.lib, tasks.nbytes); h_tasks = p_tasks.view(api.ALIGN_TASK, n); h_tasks[:] = tasks
    p_bases = api.PinnedArray(g.lib, w.reads.nbytes); h_bases = p_bases.view(np.uint8, len(w.reads)); h_bases[:] = w.reads
    p_res = api.PinnedArray(g.lib, n * api.ALIGN_RESULT.itemsize); h_res = p_res.view(api.ALIGN_RESULT, n)
    p_ops = api.PinnedArray(g.lib, cap); h_ops = p_ops.view(np.uint8, cap)
    reads_struct = api.Reads(h_bases.ctypes.data, read_off.ctypes.data, w.n_reads)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident arm (value) ----
    g.upload_reads(h_bases, read_off)
    g.upload_align_tasks(h_tasks)
    g.sync()
    for _ in range(a.warmup):
        g.run_align(); g.sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = g.stats().kernel_launches
    barrier()
    t0 = time.perf_counter()
    dev_ms, main_ms = [], []
    for _ in range(a.steps):
        g.run_align(); g.sync()
        st = g.stats()
        dev_ms.append(st.last_run_ms); main_ms.append(st.last_main_kernel_ms)
    barrier()
    wall = time.perf_counter() - t0
    st = g.stats()
    tl_start, tl_end = g.class_timeline()
    launches = int(st.kernel_launches - l0)
    main_wc = int(st.last_main_word_columns)
    # device-event time per step (prep + sort + scans + kernels, on the library's stream)
    step_ms = float(np.mean(dev_ms))

    # ---- e2e arm: host buffers in, host buffers out ----
    def e2e_step():
        rc = g.lib.lf_gpu_align_batch(g.ctx, C.byref(reads_struct), h_tasks.ctypes.data, n, h_res.ctypes.data, h_ops.ctypes.data, cap)
        if rc != 0:
            raise SystemExit(f"lf_gpu_align_batch failed: {rc} {g.lib.lf_gpu_last_error(g.ctx).decode()}")
    for _ in range(max(1, a.warmup // 2)):
        e2e_step()
    barrier()
    t1 = time.perf_counter()
    for _ in range(a.steps):
        e2e_step()
    barrier()
    e2e_wall = time.perf_counter() - t1
    # ---- chain-level e2e: lf_gpu_align_chains (the batched alignChain_edlib): chains in, Sam_t records out ----
    seeds_a, chains_a = api.workload_chains(w)
    cg = api.Contigs(w.contig_off.ctypes.data, w.contig_len.ctypes.data, len(w.contig_off))
    pac_ptr = g.pac.ctypes.data

    def chain_step():
        out = C.c_void_p()
        rc = g.lib.lf_gpu_align_chains(g.ctx, C.byref(reads_struct), C.byref(cg), seeds_a.ctypes.data, chains_a.ctypes.data, len(chains_a), pac_ptr, C.byref(out))
        if rc != 0:
            raise SystemExit(f"lf_gpu_align_chains failed: {rc} {g.lib.lf_gpu_last_error(g.ctx).decode()}")
        nrec, tbytes = C.c_size_t(), C.c_size_t()
        g.lib.lf_chain_results_records(out, C.byref(nrec))
        g.lib.lf_chain_results_text(out, C.byref(tbytes))
        cst = api.ChainStats()
        g.lib.lf_chain_results_stats(out, C.byref(cst))
        g.lib.lf_chain_results_free(out)
        return nrec.value, cst, tbytes.value
    chain_step()
    barrier()
    t2 = time.perf_counter()
    nrec, cst, chain_text_bytes = 0, None, 0
    for _ in range(max(2, a.steps // 4)):
        nrec, cst, chain_text_bytes = chain_step()
    barrier()
    chain_ms = (time.perf_counter() - t2) / max(2, a.steps // 4) * 1e3
    # ---- the same chunk as K concurrent calls on K contexts of this GPU (the ABI allows calls on distinct contexts from
    #      different host threads): uploads, kernels, emit and downloads of the sub-chunks overlap ----
    K = a.e2e_calls if a.e2e_calls > 0 else max(1, min(4, (os.cpu_count() or 4) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))) // 4))
    chainK_ms = None
    if K > 1:
        from concurrent.futures import ThreadPoolExecutor
        ctxs = [g] + [api.LfGpu(w.pac, len(w.ref)) for _ in range(K - 1)]
        subs = []
        for k in range(K):
            lo, hi = w.n_reads * k // K, w.n_reads * (k + 1) // K
            offs = np.ascontiguousarray(read_off[lo:hi + 1] - read_off[lo])
            ch = chains_a[lo:hi].copy()
            s_lo = int(ch["seed_off"][0])
            ch["seed_off"] -= s_lo
            ch["read_id"] -= lo
            rs = api.Reads(h_bases.ctypes.data + int(read_off[lo]), offs.ctypes.data, hi - lo)
            subs.append((ctxs[k], rs, offs, ch, seeds_a.ctypes.data + s_lo * seeds_a.itemsize))
        os.environ["LF_HOST_THREADS"] = str(max(2, (os.cpu_count() or 2) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))) // K))

        def sub_call(k):
            gk, rs, offs, ch, sp = subs[k]
            out = C.c_void_p()
            rc = gk.lib.lf_gpu_align_chains(gk.ctx, C.byref(rs), C.byref(cg), sp, ch.ctypes.data, len(ch), pac_ptr, C.byref(out))
            if rc != 0:
                raise SystemExit(f"lf_gpu_align_chains failed: {rc} {gk.lib.lf_gpu_last_error(gk.ctx).decode()}")
            nr, tb = C.c_size_t(), C.c_size_t()
            gk.lib.lf_chain_results_records(out, C.byref(nr))
            gk.lib.lf_chain_results_text(out, C.byref(tb))
            gk.lib.lf_chain_results_free(out)
            return nr.value, tb.value
        pool = ThreadPoolExecutor(K)
        for _ in range(2):
            resK = list(pool.map(sub_call, range(K)))
        barrier()
        t3 = time.perf_counter()
        for _ in range(max(2, a.steps // 4)):
            resK = list(pool.map(sub_call, range(K)))
        barrier()
        chainK_ms = (time.perf_counter() - t3) / max(2, a.steps // 4) * 1e3
        assert sum(r[0] for r in resK) == nrec and sum(r[1] for r in resK) == chain_text_bytes, "sub-chunk calls returned different totals"
        os.environ.pop("LF_HOST_THREADS", None)
        for gk in ctxs[1:]:
            gk.close()
    clocks = sampler.finish() if rank == 0 else None
    ops_bytes = int(h_res["ops_len"].astype(np.int64).sum() // 4)
    h2d = int(w.reads.nbytes + read_off.nbytes + tasks.nbytes)
    d2h = int(n * api.ALIGN_RESULT.itemsize + int(st_ops_words(g, h_res)) * 4)

    # max over ranks
    tt = torch.tensor([step_ms, e2e_wall / a.steps * 1e3, wall / a.steps * 1e3, chain_ms, chainK_ms if chainK_ms is not None else chain_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    bases_all = torch.tensor([float(total_bases)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(bases_all, op=dist.ReduceOp.SUM)
    step_ms_max, e2e_ms_max, wall_ms_max, chain_ms_max, chainK_ms_max = [float(x) for x in tt.tolist()]
    bases_sum = float(bases_all.item())

    if rank == 0:
        value = bases_sum / 1e6 / (wall_ms_max * 1e-3)        # Mbp/s, wall time of run+sync per step
        e2e_v = bases_sum / 1e6 / (e2e_ms_max * 1e-3)
        # roofline of the dominant kernels (rank 0's own launches)
        try:
            peaks = {name: g.int32_peak(k) for k, name in enumerate(["lop3", "iadd3", "lop3_iadd3", "lop3_imad"])}
        except Exception as e:  # pragma: no cover
            peaks = {"error": str(e)}
        peak = peaks.get("iadd3") or 18.6
        mk_ms = float(np.mean(main_ms))
        achieved = 16.0 * main_wc / (mk_ms * 1e-3) / 1e12 if mk_ms > 0 else 0.0
        cells = int(tasks["q_len"].astype(np.int64) @ tasks["t_len"].astype(np.int64))
        out = {
            "metric": "aligned Mbp/s (alignment stage)", "value": value, "unit": "Mbp/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": wall_ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": wl_name, "reads_per_gpu": w.n_reads, "read_len": WORKLOADS[wl_name][2], "tasks_per_gpu": n,
                       "cells_per_gpu": cells, "l2": "inputs+outputs of a step (>300 MB) exceed the 126 MB L2; no explicit flush",
                       "device_ms_per_step": step_ms_max},
            "gcups": cells * world / (wall_ms_max * 1e-3) / 1e9,
            # headline: the reference-facing operator (batched alignChain_edlib), host buffers in, CIGAR/MD/NM records out
            "e2e": {"value": bases_sum / 1e6 / (chainK_ms_max * 1e-3), "unit": "Mbp/s", "ms_per_step": chainK_ms_max, "calls_per_step": K,
                    "single_call": {"value": bases_sum / 1e6 / (chain_ms_max * 1e-3), "ms_per_step": chain_ms_max},
                    # in: reads + offsets + seeds + chains + 17 B of per-chain bases / guards (the round-1 tasks are generated on the device); out: CIGAR/MD text + records
                    "h2d_bytes_per_step": int(w.reads.nbytes + read_off.nbytes + seeds_a.nbytes + chains_a.nbytes) + 17 * len(chains_a), "d2h_bytes_per_step": int(chain_text_bytes) + int(nrec) * 56, "call": "lf_gpu_align_chains"},
            "e2e_align_batch": {"value": e2e_v, "unit": "Mbp/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_max,
                                "call": "lf_gpu_align_batch (round-1 tasks only: task list in, distances + 2-bit op stream out)"},
            "e2e_chains": {"value": bases_sum / 1e6 / (chain_ms_max * 1e-3), "unit": "Mbp/s", "ms_per_step": chain_ms_max, "records_per_gpu": int(nrec),
                           "host_phase_ms": {"tasks": round(cst.ms_tasks, 2), "round1": round(cst.ms_round1, 2), "rounds2_3": round(cst.ms_rounds23, 2), "emit": round(cst.ms_emit, 2), "merge": round(cst.ms_merge, 2)},
                           "rounds": {"round1_tasks": int(cst.round1_tasks), "round2_extends": int(cst.round2_extends), "round3_tasks": int(cst.round3_tasks)},
                           "what": "lf_gpu_align_chains: chains + reads from host memory in, CIGAR/MD/NM records out (3 GPU rounds + host emit)"},
            "gpu_launches": launches,
            "roofline": {"bound": "int32", "kernel": "alignment kernels of a step (k_myers_band / k_myers_bandreg / k_myers_small size classes + k_myers_large, concurrent streams)", "achieved": achieved, "peak": peak,
                         "unit": "Top/s", "frac": achieved / peak if peak else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum summed over the 23 alignment kernels of one config-2 step, one ncu --set full
                         # capture (profiles/r02i_ncu_full_alignment_kernels_final.txt): bytes per launch set, like `achieved`; 2.0 GB of it
                         # are the scattered per-task accesses of the 2.2 M tasks of q <= 128, 1.2 GB the large-task kernel plane scratch
                         "traffic": 3.82e9 if wl_name == "config2_4.6Mbp_20kx10k" and a.sv_frac == 0.10 else None, "traffic_unit": "B per launch set",
                         "algorithmic_ops_per_launch_set": 16.0 * main_wc, "kernel_ms": mk_ms,
                         "peak_source": "lf_gpu_int32_peak (IADD3 stream, 64 SASS-verified ops/iteration) measured in this run", "int32_peaks_tops": peaks},
            "clocks": clocks,
            "class_timeline_ms": {api.CLASS_NAMES[c]: [round(float(tl_start[c]), 3), round(float(tl_end[c]), 3)] for c in range(len(tl_end)) if c != 17 and tl_end[c] >= 0},
        }
        if not a.no_cpu_baseline:
            v, kind, cores, sample = cpu_reference_rate(w, os.cpu_count() or 1, max_chains=4000)
            out["cpu_baseline"] = {"value": v, "unit": "Mbp/s", "cores": cores, "kind": kind, "sample": sample}
        print(json.dumps(out))
    for p in (p_tasks, p_bases, p_res, p_ops):
        p.free()
    g.close()
    if world > 1:
        dist.destroy_process_group()


def st_ops_words(g, h_res):
    """bytes of op stream copied back per step = slot words (16 ops each)"""
    # the library copies the whole used slot range; recompute it from the results' slot layout
    return int((h_res["ops_off"].astype(np.int64) + h_res["ops_len"].astype(np.int64)).max() // 16 + 1) if len(h_res) else 0


if __name__ == "__main__":
    main()
