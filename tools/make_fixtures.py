#!/usr/bin/env python
"""Writes fixtures/<name>.npz: the chains the REFERENCE's own front-end (seeding, window selection, chaining) hands to
alignChain_edlib for a synthetic BASELINE config, and the Sam_t records the reference produced for each of them.

Build container only (needs oracle/_ref, i.e. /root/reference):

    python tools/make_fixtures.py config2 config3 config4 mini3 mini4     # or no arguments: every missing fixture

Per dataset: reference + reads from lordfast_b200.sim (seeded) -> ref.fa / reads.fa in a scratch directory ->
`oracle/_ref/lordfast --index` -> `oracle/_ref/lordfast_chaindump --search ... -n numMap` (the stock search loop with the
reference's `alignChain` hook pointed at a recorder, oracle/ref_shim.cpp) -> chains + records -> npz.  The .pac the
reference wrote is compared with sim.pack_pac so that the regenerated reference on the GPU box is the indexed one.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lordfast_b200 import fixtures, sim  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")
FULL_READS = 40   # reads whose CIGAR / MD strings are kept verbatim (debugging aid); the others keep length + CRC-32


def write_fasta(path, name, seq):
    s = seq.tobytes().decode()
    with open(path, "w") as f:
        f.write(">%s\n" % name)
        f.write("\n".join(s[i:i + 80] for i in range(0, len(s), 80)) + "\n")


def make(name, keep_tmp=False):
    p = fixtures.DATASETS[name]
    t0 = time.time()
    ref, w = fixtures.make_inputs(name)
    tmp = tempfile.mkdtemp(prefix="lffix_" + name + "_", dir=os.environ.get("LF_FIXTURE_TMP", "/tmp"))
    write_fasta(os.path.join(tmp, "ref.fa"), "chr1", ref)
    rb = w.reads.tobytes().decode()
    with open(os.path.join(tmp, "reads.fa"), "w") as f:
        for i in range(w.n_reads):
            f.write(">r%d\n%s\n" % (i, rb[w.read_off[i]:w.read_off[i + 1]]))
    del rb
    t1 = time.time()
    subprocess.check_call([os.path.join(REFDIR, "lordfast"), "--index", "ref.fa"], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t2 = time.time()
    pac = np.fromfile(os.path.join(tmp, "ref.fa.pac"), dtype=np.uint8)
    mine = sim.pack_pac(ref)
    assert np.array_equal(pac[:len(mine)], mine), "the reference's .pac differs from sim.pack_pac"
    env = dict(os.environ, LF_CHAIN_DUMP=os.path.join(tmp, "chains.txt"), LF_CHAIN_DUMP_HASH="1", LF_CHAIN_DUMP_FULL=str(FULL_READS))
    subprocess.check_call([os.path.join(REFDIR, "lordfast_chaindump"), "--search", "ref.fa", "--seq", "reads.fa", "-t", str(os.cpu_count() or 1),
                           "-n", str(p["num_map"]), "-o", "out.sam"], cwd=tmp, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t3 = time.time()
    # per read: its chains in the order the reference aligned them (one thread maps a read from start to end)
    per_read = {}
    cur = None
    for line in open(os.path.join(tmp, "chains.txt")):
        f = line.rstrip("\n").split("\t")
        if f[0] == "C":
            rid = int(f[1][1:])
            sd = np.array([int(x) for s in f[5].split(";") if s for x in s.split(",")], dtype=np.uint32).reshape(-1, 3)
            cur = dict(rev=int(f[3]), seeds=sd, recs=[], full=[])
            assert int(f[2]) == int(w.read_off[rid + 1] - w.read_off[rid])
            per_read.setdefault(rid, []).append(cur)
        elif f[0] == "H":
            cur["recs"].append(tuple(int(x) for x in f[1:11]))
        elif f[0] == "S":
            cur["full"].append({"cigar": f[7], "md": f[8]})
    seeds, seed_off, c_read, c_rev, recs, rec_chain, full = [], [0], [], [], [], [], {}
    for rid in sorted(per_read):
        for c in per_read[rid]:
            ci = len(c_read)
            seeds.append(c["seeds"]); seed_off.append(seed_off[-1] + len(c["seeds"]))
            c_read.append(rid); c_rev.append(c["rev"])
            for r in c["recs"]:
                recs.append(r); rec_chain.append(ci)
            if c["full"]:
                assert len(c["full"]) == len(c["recs"])
                full[str(ci)] = c["full"]
    recs = np.array(recs, dtype=np.int64).reshape(-1, 10)
    os.makedirs(os.path.dirname(fixtures.path(name)), exist_ok=True)
    params = dict(p, name=name, chains=len(c_read), records=len(recs), reads_mapped=len(per_read), total_bases=int(w.total_bases))
    np.savez_compressed(
        fixtures.path(name), params=np.frombuffer(json.dumps(params).encode(), dtype=np.uint8), full=np.frombuffer(json.dumps(full).encode(), dtype=np.uint8),
        ref_crc=np.uint32(zlib.crc32(ref.tobytes())), reads_crc=np.uint32(zlib.crc32(w.reads.tobytes())),
        seeds=np.concatenate(seeds), chain_seed_off=np.array(seed_off, dtype=np.int64), chain_read=np.array(c_read, dtype=np.uint32), chain_rev=np.array(c_rev, dtype=np.uint8),
        rec_chain=np.array(rec_chain, dtype=np.uint32), rec_flag=recs[:, 0].astype(np.uint32), rec_pos=recs[:, 1].astype(np.uint32), rec_posEnd=recs[:, 2].astype(np.uint32),
        rec_qStart=recs[:, 3].astype(np.uint32), rec_qEnd=recs[:, 4].astype(np.uint32), rec_nm=recs[:, 5].astype(np.int32), rec_cigar_len=recs[:, 6].astype(np.uint32),
        rec_cigar_crc=recs[:, 7].astype(np.uint32), rec_md_len=recs[:, 8].astype(np.uint32), rec_md_crc=recs[:, 9].astype(np.uint32))
    print(f"{name}: {w.n_reads} reads / {w.total_bases / 1e6:.1f} Mbp, {len(c_read)} chains ({len(c_read) / max(1, len(per_read)):.2f} per mapped read), "
          f"{len(recs)} records, {seed_off[-1]} seeds; inputs {t1 - t0:.0f} s, index {t2 - t1:.0f} s, front-end + alignChain {t3 - t2:.0f} s; "
          f"{os.path.getsize(fixtures.path(name)) / 1e6:.1f} MB", flush=True)
    if keep_tmp:
        print("scratch kept:", tmp)
    else:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or [n for n in fixtures.DATASETS if not fixtures.available(n)]
    for n in names:
        make(n, keep_tmp="--keep" in sys.argv)
