#!/usr/bin/env python
"""bench.py -- aligned Mbp/s of lordFAST's per-candidate alignment stage on B200 (BASELINE.json metric).

A step = one pass of the stage over one chunk of reads: every candidate chain the reference's own front-end (seeding,
window selection, chaining) produced for the reads of a BASELINE config goes through alignChain_edlib's work -- head /
gap / tail alignments with CIGAR paths, clip / split extensions, follow-ups, CIGAR / MD / NM records.

Workloads (--workload; config.workload in the output line):
  config2_real   configs[1]: 4.6 Mbp reference, 20 000 x 10 kbp reads at 12-15 %, 10 % SV mix; chains dumped from the
                 reference front-end (fixtures/config2.npz, tools/make_fixtures.py).  The default when the fixture exists.
  config3_real   configs[2]: 64 Mbp reference, 15 kbp reads at 15 % (10 000 of the 100 000 reads)
  config4_real   configs[3] shape: 256 Mbp reference with duplicated segments, 20 kbp reads, --numMap 10 (5 000 reads)
  config2_model  the same reads as config2_real with chains from the front-end MODEL of lordfast_b200/sim.py (round-1
                 workload; the default when no fixture travelled); config1_model: configs[0]

  value      whole-job Mbp/s with the chunk resident in HBM: k_pack_reads + k_align_prep + sort + scans + every alignment
             kernel of round 1 (the task list lf_gpu_align_chains issues: heads / tails above _pf_clipLen distance-only)
  e2e        the reference-facing operator lf_gpu_align_chains with HOST buffers: reads + seeds + chains in (H2D), CIGAR /
             MD text + records out (D2H), all rounds inside.  value = steady state with `in_flight` calls on as many
             contexts (the ABI allows concurrent calls on distinct contexts); single_call = one call at a time
  roofline   all alignment kernels of a step (they run concurrently): 16 INT32 ops per (32-row word x column), single
             pass, full matrix against the LOP3 / IADD3 issue rate measured live by lf_gpu_int32_peak;
             traffic and the per-family figures come from the committed ncu pass of the same command (profiles/)
  cpu_baseline  the reference's own alignChain_edlib (oracle/_ref) over the same chains on the host cores

`--impl reference` times only that CPU arm on the same workload.  N > 1: one process per GPU (torchrun), every rank a
chunk of the same size (weak scaling), no collective on the data path.
"""
import argparse
import csv
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
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # one hardware queue per class stream; before torch starts CUDA

MODEL_WORKLOADS = {
    # name: (ref_len, n_reads, read_len, err_lo, err_hi)
    "config1_model": (1_000_000, 200, 10_000, 0.15, 0.15),
    "config2_model": (4_600_000, 20_000, 10_000, 0.12, 0.15),
}
REAL_WORKLOADS = {"config2_real": "config2", "config3_real": "config3", "config4_real": "config4", "mini3_real": "mini3", "mini4_real": "mini4"}
NCU_CSV = os.path.join(ROOT, "profiles", "r03_ncu_step_metrics.csv")      # tools/gpu/ncu_step.sh on the default workload
SWEEP = os.path.join(ROOT, "profiles", "r03_kernel_microbench_config5.jsonl")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=list(MODEL_WORKLOADS) + list(REAL_WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sv-frac", type=float, default=0.10, help="model workloads: fraction of SV/chimera reads (0.10 = the configured mix)")
    ap.add_argument("--in-flight", type=int, default=0, help="e2e: lf_gpu_align_chains calls in flight (contexts); 0 = one per 4 host cores of this rank, at most 4")
    ap.add_argument("--e2e-calls", type=int, default=None, help=argparse.SUPPRESS)   # round-1 name of --in-flight
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


class StageInput:
    """What the stage is handed for one chunk: reads, the 2-bit reference, candidate chains (+ their round-1 tasks)."""

    def __init__(self, name, a, rank):
        from lordfast_b200 import api, fixtures, sim
        from lordfast_b200.chain_tasks import round1_tasks
        self.name = name
        if name in REAL_WORKLOADS:
            fx = fixtures.load(REAL_WORKLOADS[name])
            self.ref_len, self.pac, self.reads, self.read_off = len(fx.ref), fx.pac, fx.reads, fx.read_off
            self.seeds, self.chains = fx.seeds, fx.chains
            self.chains_from = "reference front-end (oracle/_ref/lordfast_chaindump), " + fx.params["what"]
            self.read_len = fx.params["read_len"]
            self.fixture = fx
        else:
            ref_len, n_reads, read_len, e0, e1 = MODEL_WORKLOADS[name]
            w = sim.make_workload(ref_len, n_reads, read_len, e0, e1, seed=100 + rank, sv_frac=a.sv_frac)
            self.ref_len, self.pac, self.reads, self.read_off = len(w.ref), w.pac, w.reads, w.read_off.astype(np.uint64)
            self.seeds, self.chains = api.workload_chains(w)
            self.chains_from = "front-end model (lordfast_b200/sim.py), SV fraction %.2f" % a.sv_frac
            self.read_len = read_len
            self.fixture = None
        self.n_reads = len(self.read_off) - 1
        self.total_bases = int(self.read_off[-1])
        self.contig_off, self.contig_len = np.array([0], dtype=np.int64), np.array([self.ref_len], dtype=np.int32)
        s3 = np.stack([self.seeds["tPos"], self.seeds["qPos"], self.seeds["len"]], axis=1)
        seed_off = np.concatenate([self.chains["seed_off"], [len(self.seeds)]]).astype(np.int64)
        rl = np.diff(self.read_off.astype(np.int64))[self.chains["read_id"]]
        self.tasks, _, _ = round1_tasks(s3, seed_off, self.chains["is_rev"].astype(np.uint8), rl, self.contig_off, self.contig_len,
                                        read_id=self.chains["read_id"].astype(np.uint32))


def default_workload():
    from lordfast_b200 import fixtures
    return "config2_real" if fixtures.available("config2") else "config2_model"


def cpu_reference_rate(si, nthreads, max_chains=None):
    """Mbp/s of the reference's alignChain_edlib (oracle/_ref) over this workload's chains (reads counted once)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    from lordfast_b200 import sim
    n = len(si.chains) if max_chains is None else min(len(si.chains), max_chains)
    ch = si.chains[:n]
    rids = ch["read_id"].astype(np.int64)
    ro = si.read_off.astype(np.int64)
    bases = int(sum(int(ro[r + 1] - ro[r]) for r in np.unique(rids)))
    if O.have_ref():
        lib = O.ref()
        idx_off = (C.c_int64 * 1)(0)
        idx_len = (C.c_int32 * 1)(si.ref_len)
        lib.ref_set_index(si.pac.ctypes.data, si.ref_len, 1, idx_off, idx_len)
        oriented = []
        for c in ch:
            r = si.reads[ro[c["read_id"]]:ro[c["read_id"] + 1]]
            oriented.append(np.ascontiguousarray(sim.revcomp(r) if c["is_rev"] else r).tobytes())
        qptr = (C.c_char_p * n)(*oriented)
        s_hi = int(ch["seed_off"][-1] + ch["n_seeds"][-1])
        seeds = np.ascontiguousarray(np.stack([si.seeds["tPos"][:s_hi], si.seeds["qPos"][:s_hi], si.seeds["len"][:s_hi]], axis=1), dtype=np.uint32)
        seed_off = np.ascontiguousarray(np.concatenate([ch["seed_off"], [s_hi]]), dtype=np.int64)
        rl = np.array([len(o) for o in oriented], dtype=np.int32)
        isrev = np.ascontiguousarray(ch["is_rev"], dtype=np.uint8)
        nsam = C.c_int64()
        lib.ref_replay_chains.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        lib.ref_replay_chains.restype = C.c_double
        secs = lib.ref_replay_chains(n, seed_off.ctypes.data, seeds.ctypes.data, C.cast(qptr, C.c_void_p), rl.ctypes.data,
                                     isrev.ctypes.data, nthreads, C.byref(nsam))
        return bases / secs / 1e6, "reference", nthreads, f"{n} chains / {bases / 1e6:.1f} Mbp of reads through the reference's alignChain_edlib (oracle/_ref), {nthreads} threads"
    # oracle/_ref absent: the oracle port, single thread, small sample
    n = min(n, 100)
    idx = O.PacIndex(si.pac, si.ref_len, si.contig_off, si.contig_len)
    t0 = time.time()
    seen, bases = set(), 0
    for c in ch[:n]:
        rid = int(c["read_id"])
        r = si.reads[ro[rid]:ro[rid + 1]]
        q = (sim.revcomp(r) if c["is_rev"] else r).tobytes()
        sd = si.seeds[int(c["seed_off"]):int(c["seed_off"]) + int(c["n_seeds"])]
        O.oracle_align_chain(idx, [(int(x["tPos"]), int(x["qPos"]), int(x["len"])) for x in sd], q, int(c["is_rev"]))
        if rid not in seen:
            seen.add(rid); bases += len(q)
    secs = time.time() - t0
    return bases / secs / 1e6, "port", 1, f"{n} chains / {bases / 1e6:.1f} Mbp through oracle/lf_oracle.c (scalar O(q*t) port), 1 thread"


def ncu_figures(counts):
    """Per-kernel figures of the committed ncu metrics pass (same workload, one resident step): DRAM bytes per step, time
    of each kernel family run alone, warp instructions.  None if the file is absent."""
    if not os.path.exists(NCU_CSV):
        return None
    rows = list(csv.reader(open(NCU_CSV)))
    hdr, per, prep_ids = None, {}, []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(LfDev")[0].replace("void ", "")
        if d["Metric Name"] == "gpu__time_duration.sum" and name.startswith("k_align_prep"):
            prep_ids.append(int(d["ID"]))
        if not name.startswith("k_myers"):
            continue
        key = (name, d["ID"])
        v = float(d["Metric Value"].replace(",", ""))
        if d["Metric Name"] == "gpu__time_duration.sum":
            u = d["Metric Unit"]
            v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
        per.setdefault(key, {})[d["Metric Name"]] = v
    # the first resident step = the launches before any kernel name repeats beyond what one step holds
    fam = {"q<=128 (k_myers_bandreg<1..4,false>, k_myers_small<1..4,shw>)": ["bandreg<1, 0>", "bandreg<2, 0>", "bandreg<3, 0>", "bandreg<4, 0>", "small<1,", "small<2,", "small<3,", "small<4,"],
           "129..512 rows near the diagonal (k_myers_bandreg<3..5,true>)": ["bandreg<3, 1>", "bandreg<4, 1>", "bandreg<5, 1>"],
           "129..512 rows, other (k_myers_small<6..16>, k_myers_group<4,4>)": ["small<6,", "small<8,", "small<12,", "small<16,", "group<4, 4", "group<16, 1"],
           "above 512 rows (k_myers_group, k_myers_large)": ["group<8,", "group<16, 8", "group<32,", "k_myers_large"]}
    # one resident step = the alignment kernels between the first k_align_prep and the next one
    prep_ids.sort()
    lo = prep_ids[0] if prep_ids else -1
    hi = prep_ids[1] if len(prep_ids) > 1 else 1 << 30
    out = {"traffic": 0.0, "families": {}, "warp_instructions": 0.0}
    step_ids = [k for k in per if lo < int(k[1]) < hi]
    for key in step_ids:
        m = per[key]
        out["traffic"] += m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
        out["warp_instructions"] += m.get("smsp__inst_executed.sum", 0)
        for f, pats in fam.items():
            if any(p in key[0] for p in pats):
                e = out["families"].setdefault(f, {"alone_ms": 0.0, "warp_instructions": 0.0, "dram_bytes": 0.0})
                e["alone_ms"] += m.get("gpu__time_duration.sum", 0)
                e["warp_instructions"] += m.get("smsp__inst_executed.sum", 0)
                e["dram_bytes"] += m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
                break
    return out


def family_word_columns(tasks):
    """algorithmic word-columns (ceil(q/32) * t) of the round-1 tasks per kernel family, by the routing rules of k_align_prep"""
    q, t = tasks["q_len"].astype(np.int64), tasks["t_len"].astype(np.int64)
    nw = (q + 31) // 32
    wc = nw * t
    dlt = np.abs(q - t)
    sc_nb = np.where(nw <= 6, 3, np.where(nw <= 12, 4, 5))
    near = (tasks["mode"] == 0) & (3 * dlt <= 32 * (sc_nb - 1) - 7) & (q // np.maximum(t, 1) < 32)
    leaf = (20 * ((q + 63) // 64) * t + 8 * t < (1 << 20)) | (t < 2)
    small = (nw <= 4) & leaf
    mid = (nw > 4) & (nw <= 16) & leaf
    return {"q<=128 (k_myers_bandreg<1..4,false>, k_myers_small<1..4,shw>)": int(wc[small].sum()),
            "129..512 rows near the diagonal (k_myers_bandreg<3..5,true>)": int(wc[mid & near].sum()),
            "129..512 rows, other (k_myers_small<6..16>, k_myers_group<4,4>)": int(wc[mid & ~near].sum()),
            "above 512 rows (k_myers_group, k_myers_large)": int(wc[~small & ~mid].sum())}


def sweep_summary():
    if not os.path.exists(SWEEP):
        return None
    pts = [json.loads(l) for l in open(SWEEP) if l.strip().startswith("{")]
    nw = [p for p in pts if p.get("kind") in ("NW", "SHW") and "roofline_frac" in p]
    if not nw:
        return None
    lo, hi = min(nw, key=lambda p: p["roofline_frac"]), max(nw, key=lambda p: p["roofline_frac"])
    brief = lambda p: {k: p[k] for k in ("kind", "L", "div", "pairs", "kernel_ms", "gcups", "roofline_frac", "oracle_mismatches") if k in p}
    return {"source": os.path.relpath(SWEEP, ROOT), "points": len(nw), "worst": brief(lo), "best": brief(hi)}


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl_name = a.workload or default_workload()
    if a.e2e_calls is not None:
        a.in_flight = a.e2e_calls

    if a.impl == "reference":
        if rank != 0:
            return
        si = StageInput(wl_name, a, 0)
        nthreads = os.cpu_count() or 1
        # each step = the whole chunk's chains (the same ones the GPU arm aligns) unless that takes over ~8 s
        r0, kind, cores, sample = cpu_reference_rate(si, nthreads, max_chains=2000)
        full_s = si.total_bases / 1e6 / r0
        max_chains = None if full_s * (a.steps + a.warmup) < 200 else max(2000, int(len(si.chains) * 200 / (full_s * (a.steps + a.warmup))))
        rates = []
        for it in range(a.warmup + a.steps):
            r, kind, cores, sample = cpu_reference_rate(si, nthreads, max_chains=max_chains)
            if it >= a.warmup:
                rates.append(r)
        v = float(np.mean(rates))
        ms = si.total_bases / 1e6 / v * 1e3
        print(json.dumps({"impl": "reference", "metric": "aligned Mbp/s (alignment stage)", "value": v, "unit": "Mbp/s", "n_gpus": a.gpus,
                          "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                          "config": {"workload": wl_name, "chains": si.chains_from, "reads_per_gpu": si.n_reads, "read_len": si.read_len, "sample": sample},
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

    si = StageInput(wl_name, a, rank)
    tasks = si.tasks
    g = api.LfGpu(si.pac, si.ref_len)
    n = len(tasks)
    total_bases = si.total_bases
    read_off = si.read_off.astype(np.uint64)

    p_tasks = api.PinnedArray(g.lib, tasks.nbytes); h_tasks = p_tasks.view(api.ALIGN_TASK, n); h_tasks[:] = tasks
    p_bases = api.PinnedArray(g.lib, si.reads.nbytes); h_bases = p_bases.view(np.uint8, len(si.reads)); h_bases[:] = si.reads
    reads_struct = api.Reads(h_bases.ctypes.data, read_off.ctypes.data, si.n_reads)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident arm (value): bit planes + round-1 alignment of the chunk, inputs in HBM ----
    g.upload_reads(h_bases, read_off)
    g.upload_align_tasks(h_tasks)
    g.sync()
    for _ in range(max(3, a.warmup)):
        g.pack_reads(); g.run_align(); g.sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = g.stats().kernel_launches
    barrier()
    t0 = time.perf_counter()
    main_ms = []
    for _ in range(a.steps):
        g.pack_reads(); g.run_align(); g.sync()
        main_ms.append(g.stats().last_main_kernel_ms)
    barrier()
    wall = time.perf_counter() - t0
    st = g.stats()
    tl_start, tl_end = g.class_timeline()
    class_counts = g.class_counts()
    launches = int(st.kernel_launches - l0)
    main_wc = int(st.last_main_word_columns)

    # ---- e2e: lf_gpu_align_chains (the batched alignChain_edlib), host buffers in, records + text out ----
    cg = api.Contigs(si.contig_off.ctypes.data, si.contig_len.ctypes.data, len(si.contig_off))
    # seeds and chains in pinned memory, as the glue keeps them (integration/lordfast_gpu_glue.cpp): the library reads them as they are
    p_seeds = api.PinnedArray(g.lib, si.seeds.nbytes + 64); seeds_a = p_seeds.view(api.SEED, len(si.seeds)); seeds_a[:] = si.seeds
    p_chains = api.PinnedArray(g.lib, si.chains.nbytes + 64); chains_a = p_chains.view(api.CHAIN, len(si.chains)); chains_a[:] = si.chains

    def chain_call(gk):
        out = C.c_void_p()
        rc = gk.lib.lf_gpu_align_chains(gk.ctx, C.byref(reads_struct), C.byref(cg), seeds_a.ctypes.data, chains_a.ctypes.data, len(chains_a), gk.pac.ctypes.data, C.byref(out))
        if rc != 0:
            raise SystemExit(f"lf_gpu_align_chains failed: {rc} {gk.lib.lf_gpu_last_error(gk.ctx).decode()}")
        nrec, tbytes = C.c_size_t(), C.c_size_t()
        gk.lib.lf_chain_results_records(out, C.byref(nrec))
        gk.lib.lf_chain_results_text(out, C.byref(tbytes))
        cst = api.ChainStats()
        gk.lib.lf_chain_results_stats(out, C.byref(cst))
        gk.lib.lf_chain_results_free(out)
        return nrec.value, cst, tbytes.value

    nsingle = max(3, a.steps // 3)
    chain_call(g); chain_call(g)
    barrier()
    t2 = time.perf_counter()
    for _ in range(nsingle):
        nrec, cst, chain_text_bytes = chain_call(g)
    barrier()
    single_ms = (time.perf_counter() - t2) / nsingle * 1e3
    cores_here = (os.cpu_count() or 4) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    P = a.in_flight if a.in_flight > 0 else (4 if cores_here >= 16 else 2 if cores_here >= 4 else 1)
    flight_ms = single_ms
    tried = []
    for Pk in ([P] if (a.in_flight > 0 or P < 4) else [4, 2]):   # auto: four calls in flight, then two; the better one counts
        if Pk <= 1:
            continue
        ctxs = [g] + [api.LfGpu(si.pac, si.ref_len) for _ in range(Pk - 1)]
        os.environ["LF_HOST_THREADS"] = str(max(2, cores_here // Pk))
        per = max(2, (a.steps + Pk - 1) // Pk)

        def worker(gk, k):
            for _ in range(k):
                chain_call(gk)
        for gk in ctxs[1:]:
            chain_call(gk); chain_call(gk)
        barrier()
        t3 = time.perf_counter()
        th = [threading.Thread(target=worker, args=(gk, per)) for gk in ctxs]
        [t.start() for t in th]
        [t.join() for t in th]
        barrier()
        ms_k = (time.perf_counter() - t3) / (Pk * per) * 1e3
        tried.append((Pk, ms_k))
        os.environ.pop("LF_HOST_THREADS", None)
        for gk in ctxs[1:]:
            gk.close()
    if tried:
        if world > 1:   # every rank must pick the same setting: the one with the best worst rank
            tv = torch.tensor([m for _, m in tried], device="cuda", dtype=torch.float64)
            dist.all_reduce(tv, op=dist.ReduceOp.MAX)
            tried = [(pk, float(m)) for (pk, _), m in zip(tried, tv.tolist())]
        P, flight_ms = min(tried, key=lambda x: x[1])
    clocks = sampler.finish() if rank == 0 else None

    # max over ranks
    tt = torch.tensor([wall / a.steps * 1e3, single_ms, flight_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    bases_all = torch.tensor([float(total_bases)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(bases_all, op=dist.ReduceOp.SUM)
    wall_ms_max, single_ms_max, flight_ms_max = [float(x) for x in tt.tolist()]
    bases_sum = float(bases_all.item())

    if rank == 0:
        value = bases_sum / 1e6 / (wall_ms_max * 1e-3)
        try:
            peaks = {name: g.int32_peak(k) for k, name in enumerate(["lop3", "iadd3", "lop3_iadd3", "lop3_imad"])}
        except Exception as e:  # pragma: no cover
            peaks = {"error": str(e)}
        peak = peaks.get("iadd3") or 18.6
        mk_ms = float(np.mean(main_ms))
        achieved = 16.0 * main_wc / (mk_ms * 1e-3) / 1e12 if mk_ms > 0 else 0.0
        cells = int(tasks["q_len"].astype(np.int64) @ tasks["t_len"].astype(np.int64))
        roof = {"bound": "int32", "kernel": "alignment kernels of a step (k_myers_bandreg / k_myers_small size classes, k_myers_group, k_myers_large; concurrent streams)",
                "achieved": achieved, "peak": peak, "unit": "Top/s", "frac": achieved / peak if peak else None, "traffic": None, "traffic_unit": "B per launch set",
                "algorithmic_ops_per_launch_set": 16.0 * main_wc, "kernel_ms": mk_ms,
                "peak_source": "lf_gpu_int32_peak (IADD3 stream, 64 SASS-verified ops/iteration) measured in this run", "int32_peaks_tops": peaks}
        nf = ncu_figures(class_counts) if wl_name == "config2_real" else None
        if nf:
            roof["traffic"] = nf["traffic"]
            roof["traffic_source"] = os.path.relpath(NCU_CSV, ROOT) + " (dram__bytes_read.sum + dram__bytes_write.sum over the alignment kernels of one step, ncu --metrics pass of this command)"
            roof["executed_warp_instructions"] = nf["warp_instructions"]
            fw = family_word_columns(tasks)
            roof["families"] = {f: {"word_columns": fw.get(f, 0), "alone_ms": round(e["alone_ms"], 4), "warp_instructions": e["warp_instructions"], "dram_bytes": e["dram_bytes"],
                                    "frac_alone": (16.0 * fw.get(f, 0) / (e["alone_ms"] * 1e-3) / 1e12 / peak) if e["alone_ms"] > 0 and peak else None,
                                    "executed_per_algorithmic_instruction": e["warp_instructions"] * 32 / (16.0 * fw[f]) if fw.get(f) else None}
                                for f, e in nf["families"].items()}
            roof["families_note"] = "alone_ms: the family's kernels timed one at a time under ncu (cold, serialised); in a step they overlap, so the fractions do not add up to `frac`"
        sw = sweep_summary()
        out = {
            "metric": "aligned Mbp/s (alignment stage)", "value": value, "unit": "Mbp/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": wall_ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": wl_name, "chains": si.chains_from, "reads_per_gpu": si.n_reads, "read_len": si.read_len, "chains_per_gpu": int(len(si.chains)),
                       "tasks_per_gpu": n, "cells_per_gpu": cells, "word_columns_per_gpu": main_wc,
                       "step": "k_pack_reads + k_align_prep + sort + scans + all round-1 alignment kernels, chunk resident in HBM",
                       "l2": "inputs+outputs of a step (>300 MB) exceed the 126 MB L2; no explicit flush"},
            "gcups": cells * world / (wall_ms_max * 1e-3) / 1e9,
            # headline: the reference-facing operator (batched alignChain_edlib), host buffers in, CIGAR/MD/NM records out
            "e2e": {"value": bases_sum / 1e6 / (min(flight_ms_max, single_ms_max) * 1e-3), "unit": "Mbp/s", "ms_per_step": min(flight_ms_max, single_ms_max),
                    "calls_in_flight": P if flight_ms_max < single_ms_max else 1, "in_flight_tried": {"calls": P, "ms_per_step": flight_ms_max, "all": [{"calls": pk, "ms_per_step": m} for pk, m in tried]},
                    "single_call": {"value": bases_sum / 1e6 / (single_ms_max * 1e-3), "ms_per_step": single_ms_max},
                    # in: reads + offsets + seeds + chains + 17 B of per-chain bases / guards (the round-1 tasks are generated on the device); out: CIGAR/MD text + records
                    "h2d_bytes_per_step": int(si.reads.nbytes + read_off.nbytes + seeds_a.nbytes + chains_a.nbytes) + 17 * len(chains_a),
                    "d2h_bytes_per_step": int(chain_text_bytes) + int(nrec) * 56, "call": "lf_gpu_align_chains",
                    "records_per_gpu": int(nrec),
                    "host_phase_ms": {"tasks": round(cst.ms_tasks, 2), "round1": round(cst.ms_round1, 2), "rounds2_3": round(cst.ms_rounds23, 2), "emit": round(cst.ms_emit, 2)},
                    "rounds": {"round1_tasks": int(cst.round1_tasks), "round2_extends": int(cst.round2_extends), "round3_tasks": int(cst.round3_tasks)}},
            "gpu_launches": launches,
            "roofline": roof,
            "clocks": clocks,
            "class_tasks": class_counts,
            "class_timeline_ms": {api.CLASS_NAMES[c]: [round(float(tl_start[c]), 3), round(float(tl_end[c]), 3)] for c in range(len(tl_end)) if c != 17 and tl_end[c] >= 0},
        }
        if sw:
            out["config5_sweep"] = sw
        if not a.no_cpu_baseline:
            v, kind, cores, sample = cpu_reference_rate(si, os.cpu_count() or 1, max_chains=4000)
            out["cpu_baseline"] = {"value": v, "unit": "Mbp/s", "cores": cores, "kind": kind, "sample": sample}
    p_tasks.free(); p_bases.free(); p_seeds.free(); p_chains.free()
    g.close()
    if rank == 0:
        if world == 1 and wl_name == "config2_real" and not os.environ.get("LF_BENCH_NO_SEEDING"):
            # the stage in front of this one, on the same reads: key figures of bench_seed.py (its own line has the rest)
            try:
                import bench_seed
                sd = bench_seed.measure(steps=5, warmup=3, cpu_sample=200, cpu_baseline=not a.no_cpu_baseline,
                                        data=(si.fixture.ref, np.ascontiguousarray(si.reads, dtype=np.uint8), si.read_off.astype(np.uint64), "config2_real reads (fixtures/config2.npz)"))
                out["seeding"] = {"metric": sd["metric"], "value": sd["value"], "unit": sd["unit"], "ms_per_step": sd["ms_per_step"], "e2e": sd["e2e"],
                                  "kernels": sd["kernels"], "roofline": sd["roofline"], "cpu_baseline": sd.get("cpu_baseline"), "call": "lf_gpu_seed_batch"}
            except Exception as e:   # noqa: BLE001 -- the alignment line must not depend on it
                out["seeding"] = {"unavailable": repr(e)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
