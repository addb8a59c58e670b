#!/usr/bin/env python
"""End-to-end `lordfast --search -t N`: the unmodified reference (oracle/_ref/lordfast, CPU) beside the same program
with seeding and the alignment stage on the GPU (integration/_build/lordfast_gpu; LF_GPU_SEED=0 in the environment: alignment stage only), on the same synthetic input, on this box.
Wall time = sum of the program's own "mapping... done in" lines; the SAM files are compared (sorted).
Prints one JSON line.  Nothing here reads /root/reference.

    python tools/bench_lordfast_e2e.py --ref-len 4600000 --reads 20000 --read-len 10000 --err 0.12 0.15 [-t N]
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "lordfast")
GPU_BIN = os.path.join(ROOT, "integration", "_build", "lordfast_gpu")


def run(binary, tmp, out, threads, extra):
    t0 = time.time()
    p = subprocess.run([binary, "--search", "ref.fa", "--seq", "reads.fa", "-t", str(threads), "-o", out, *extra], cwd=tmp,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, check=True)
    wall = time.time() - t0
    mapping = sum(float(x) for x in re.findall(r"done in ([0-9.]+) seconds", p.stderr))
    phases = [tuple(float(v) for v in m) for m in re.findall(r"front-end ([0-9.]+) ms, GPU alignment stage ([0-9.]+) ms, scoring\+SAM ([0-9.]+) ms", p.stderr)]
    seeding = [float(v) for v in re.findall(r"gather \+ GPU seeding ([0-9.]+) ms", p.stderr)]
    init = re.findall(r"lf_gpu_init ([0-9.]+) ms", p.stderr)
    return wall, mapping, phases, float(init[0]) if init else 0.0, seeding


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-len", type=int, default=4_600_000)
    ap.add_argument("--reads", type=int, default=20_000)
    ap.add_argument("--read-len", type=int, default=10_000)
    ap.add_argument("--err", type=float, nargs=2, default=(0.12, 0.15))
    ap.add_argument("--sv-frac", type=float, default=0.10)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("-t", "--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--replicate", type=int, default=1, help="write the read set this many times (renamed copies): longer runs, several chunks")
    ap.add_argument("extra", nargs="*")
    a = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="lfe2e")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "integration", "make_dataset.py"), tmp, "--ref-len", str(a.ref_len), "--reads", str(a.reads),
                           "--read-len", str(a.read_len), "--err", str(a.err[0]), str(a.err[1]), "--sv-frac", str(a.sv_frac), "--seed", str(a.seed)],
                          stdout=subprocess.DEVNULL)
    if a.replicate > 1:
        body = open(os.path.join(tmp, "reads.fa")).read()
        with open(os.path.join(tmp, "reads.fa"), "w") as f:
            for k in range(a.replicate):
                f.write(body.replace(">r", ">c%d_r" % k))
    bases = sum(len(l) - 1 for l in open(os.path.join(tmp, "reads.fa")) if not l.startswith(">"))
    t0 = time.time()
    subprocess.check_call([REF_BIN, "--index", "ref.fa"], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t_index = time.time() - t0
    best = {}
    for name, binary, out in (("cpu", REF_BIN, "cpu.sam"), ("gpu", GPU_BIN, "gpu.sam")):
        for _ in range(a.repeat):
            r = run(binary, tmp, out, a.threads, a.extra)
            if name not in best or r[1] < best[name][1]:
                best[name] = r
    sa = sorted(l for l in open(os.path.join(tmp, "cpu.sam")) if not l.startswith("@PG"))
    sb = sorted(l for l in open(os.path.join(tmp, "gpu.sam")) if not l.startswith("@PG"))
    ph = best["gpu"][2]
    print(json.dumps({
        "what": "lordfast --search -t N end to end: reference CPU binary vs the same program with the alignment stage on the GPU",
        "workload": {"ref_len": a.ref_len, "reads": a.reads * a.replicate, "replicate": a.replicate, "read_len": a.read_len, "err": list(a.err), "sv_frac": a.sv_frac, "read_bases": bases},
        "threads": a.threads, "host_cores": os.cpu_count(), "index_s": round(t_index, 2),
        "cpu": {"mapping_s": best["cpu"][1], "wall_s": round(best["cpu"][0], 2), "mbp_per_s": round(bases / 1e6 / best["cpu"][1], 2)},
        "gpu": {"mapping_s": best["gpu"][1], "wall_s": round(best["gpu"][0], 2), "mbp_per_s": round(bases / 1e6 / best["gpu"][1], 2),
                "chunks": len(ph), "lf_gpu_init_ms": best["gpu"][3],
                "gpu_stage_ms_per_chunk": [p[1] for p in ph],
                "phases_ms": {"gather_and_gpu_seeding": sum(best["gpu"][4]), "front_end_cpu": sum(p[0] for p in ph), "gpu_alignment_stage": sum(p[1] for p in ph), "scoring_and_sam_cpu": sum(p[2] for p in ph)}},
        "speedup_mapping": round(best["cpu"][1] / best["gpu"][1], 2),
        "sam_records": len(sa), "sam_identical_sorted": sa == sb,
    }))
    if sa != sb:
        sys.exit("SAM differs")


if __name__ == "__main__":
    main()
