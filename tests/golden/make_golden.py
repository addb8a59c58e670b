#!/usr/bin/env python
"""Generates the committed golden vectors from the REFERENCE's own code (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

  align_pairs.json   edlibAlign(k=-1, NW|SHW, PATH) -> editDistance, endLocations[0], alignment[]
  ksw_cases.json     ksw_extend2 with lordFAST's two parameter sets -> score, qle, tle
  chains.npz/.json   reads + chains dumped from the reference front-end through its alignChain hook
                     (oracle/_ref/lordfast_chaindump) and the Sam_t records alignChain_edlib produced
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import _oracle as O  # noqa: E402
from lordfast_b200 import sim  # noqa: E402


def rnd(rng, n):
    return sim.ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def main():
    assert O.have_ref(), "oracle/_ref missing: run `make -C oracle ref` first"
    rng = np.random.default_rng(20241017)
    pairs = []

    def add_pair(q, t, mode):
        q, t = bytes(q), bytes(t)
        ed, end, ops = O.ref_align(q, t, mode)
        pairs.append({"q": q.decode(), "t": t.decode(), "mode": mode, "ed": ed, "end": end, "ops": "".join(str(c) for c in ops)})

    lens = [1, 2, 3, 7, 12, 31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 129, 191, 192, 193, 255, 256, 257, 383, 384, 385, 511, 512, 513, 700, 1100]
    for L in lens:
        t = rnd(rng, L)
        for d in (0.05, 0.15, 0.30):
            q = sim.mutate_pair(t, d, rng)
            add_pair(q.tobytes(), t.tobytes(), 0)
            add_pair(q.tobytes(), np.concatenate([t, rnd(rng, 20)]).tobytes(), 1)
        add_pair(rnd(rng, max(1, L // 2)).tobytes(), t.tobytes(), 0)          # unrelated, skewed
        add_pair(rnd(rng, L).tobytes(), rnd(rng, max(1, L // 3)).tobytes(), 1)
    add_pair(b"ACGT", b"TTTT", 1)      # SHW whose best prefix is the empty one (end = -1)
    add_pair(b"A", b"C", 1)
    add_pair(b"AAAAAAAA", b"A", 0)
    add_pair(b"A", b"AAAAAAAA", 0)
    add_pair(b"ACGTNNacgt", b"ACGTACGTAC", 0)  # N and lower case never match the reference
    # sizes straddling edlib's 1 MiB rule (square switch point 1792/1793) and beyond (Hirschberg)
    for L in (1792, 1793, 2300, 3600):
        t = rnd(rng, L)
        add_pair(sim.mutate_pair(t, 0.15, rng).tobytes(), t.tobytes(), 0)
        add_pair(sim.mutate_pair(t, 0.12, rng).tobytes(), np.concatenate([t, rnd(rng, 20)]).tobytes(), 1)
    add_pair(rnd(rng, 40).tobytes(), rnd(rng, 38000).tobytes(), 0)            # one block, very long target
    add_pair(np.concatenate([sim.mutate_pair(rnd(rng, 900), 0.1, rng), rnd(rng, 1500)]).tobytes(), rnd(rng, 2420).tobytes(), 1)
    json.dump(pairs, open(os.path.join(HERE, "align_pairs.json"), "w"), separators=(",", ":"))

    code = np.zeros(256, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        code[ch] = i
    ksw = []
    for L in (5, 40, 120, 300, 600, 900, 1500, 2500):
        for kind in range(4):
            t = rnd(rng, L)
            if kind == 0:
                q = np.concatenate([sim.mutate_pair(t[: L // 2], 0.15, rng), rnd(rng, L // 2 + 1)])
            elif kind == 1:
                q = sim.mutate_pair(t, 0.2, rng)
            elif kind == 2:
                q = rnd(rng, int(rng.integers(1, L + 50)))
            else:
                q = np.concatenate([sim.mutate_pair(t[: L // 3], 0.1, rng), rnd(rng, 100), sim.mutate_pair(t[L // 3:], 0.1, rng)])
            for prm in ((0, 1, 0, 1, 40, 40), (8, 1, 4, 1, 100, 200)):
                sc, qle, tle = O.ref_extend(code[q].tobytes(), code[t].tobytes(), *prm)
                ksw.append({"q": q.tobytes().decode(), "t": t.tobytes().decode(), "prm": list(prm), "score": sc, "qle": qle, "tle": tle})
    json.dump(ksw, open(os.path.join(HERE, "ksw_cases.json"), "w"), separators=(",", ":"))

    # chains from the reference front-end
    w = sim.make_workload(300_000, 40, 4000, 0.12, 0.15, seed=5, sv_frac=0.5)
    tmp = tempfile.mkdtemp(prefix="lfgold")
    with open(os.path.join(tmp, "ref.fa"), "w") as f:
        s = w.ref.tobytes().decode()
        f.write(">chr1\n" + "\n".join(s[i:i + 80] for i in range(0, len(s), 80)) + "\n")
    with open(os.path.join(tmp, "reads.fa"), "w") as f:
        for i in range(w.n_reads):
            f.write(">r%d\n%s\n" % (i, w.reads[w.read_off[i]:w.read_off[i + 1]].tobytes().decode()))
    refdir = os.path.join(ROOT, "oracle", "_ref")
    subprocess.check_call([os.path.join(refdir, "lordfast"), "--index", "ref.fa"], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    env = dict(os.environ, LF_CHAIN_DUMP=os.path.join(tmp, "chains.txt"))
    subprocess.check_call([os.path.join(refdir, "lordfast_chaindump"), "--search", "ref.fa", "--seq", "reads.fa", "-t", "1", "-o", "out.sam"],
                          cwd=tmp, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    chains = []
    for line in open(os.path.join(tmp, "chains.txt")):
        f = line.rstrip("\n").split("\t")
        if f[0] == "C":
            seeds = [[int(x) for x in s.split(",")] for s in f[5].split(";") if s]
            chains.append({"read": int(f[1][1:]), "readLen": int(f[2]), "isRev": int(f[3]), "seeds": seeds, "sam": []})
        else:
            chains[-1]["sam"].append({"flag": int(f[1]), "pos": int(f[2]), "posEnd": int(f[3]), "qStart": int(f[4]), "qEnd": int(f[5]),
                                      "nm": int(f[6]), "cigar": f[7], "md": f[8]})
    pac = np.fromfile(os.path.join(tmp, "ref.fa.pac"), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "chains.npz"), pac=pac, l_pac=np.int64(len(w.ref)), reads=w.reads, read_off=w.read_off)
    json.dump(chains, open(os.path.join(HERE, "chains.json"), "w"), separators=(",", ":"))
    sam = [l for l in open(os.path.join(tmp, "out.sam")) if not l.startswith("@PG")]
    open(os.path.join(HERE, "chains_reference.sam"), "w").writelines(sorted(sam))
    print(f"{len(pairs)} align pairs, {len(ksw)} ksw cases, {len(chains)} chains "
          f"({sum(len(c['sam']) for c in chains)} records, {sum(1 for c in chains if len(c['sam']) > 1)} split)")


if __name__ == "__main__":
    main()
