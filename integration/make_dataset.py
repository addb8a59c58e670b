#!/usr/bin/env python
"""Writes a synthetic lordFAST input set: ref.fa (+ optional duplicated segments so that several
windows tie and the fine mode / --numMap path is taken) and reads.fa, from lordfast_b200.sim.

    python integration/make_dataset.py OUTDIR --ref-len 1000000 --reads 200 --read-len 10000 --err 0.15 0.15 [--dups 20]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lordfast_b200 import sim  # noqa: E402


def make_ref(ref_len, seed, dups, n_contigs=1):
    return sim.make_reference_dups(ref_len, seed, dups)


def write(outdir, ref, w, n_contigs=1):
    os.makedirs(outdir, exist_ok=True)
    s = ref.tobytes().decode()
    with open(os.path.join(outdir, "ref.fa"), "w") as f:
        bounds = [len(s) * i // n_contigs for i in range(n_contigs + 1)]
        for c in range(n_contigs):
            f.write(">chr%d\n" % (c + 1))
            seg = s[bounds[c]:bounds[c + 1]]
            f.write("\n".join(seg[i:i + 80] for i in range(0, len(seg), 80)) + "\n")
    rb = w.reads.tobytes().decode()
    with open(os.path.join(outdir, "reads.fa"), "w") as f:
        for i in range(w.n_reads):
            f.write(">r%d\n%s\n" % (i, rb[w.read_off[i]:w.read_off[i + 1]]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("outdir")
    ap.add_argument("--ref-len", type=int, default=1_000_000)
    ap.add_argument("--reads", type=int, default=200)
    ap.add_argument("--read-len", type=int, default=10_000)
    ap.add_argument("--err", type=float, nargs=2, default=(0.15, 0.15))
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--sv-frac", type=float, default=0.10)
    ap.add_argument("--dups", type=int, default=0)
    ap.add_argument("--contigs", type=int, default=1)
    a = ap.parse_args()
    ref = make_ref(a.ref_len, a.seed, a.dups)
    w = sim.make_workload(a.ref_len, a.reads, a.read_len, a.err[0], a.err[1], seed=a.seed, sv_frac=a.sv_frac, ref=ref)
    write(a.outdir, ref, w, a.contigs)
    print("wrote %s: %d bp reference (%d contigs, %d duplicated segments), %d reads / %d bases" %
          (a.outdir, len(ref), a.contigs, a.dups, w.n_reads, w.total_bases))


if __name__ == "__main__":
    main()
