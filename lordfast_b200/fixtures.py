"""Pre-dumped front-end chains of the reference (the stage is timed and checked on chains the
reference's own seeding / window selection / chaining produced, not on a model of them).

A fixture is written in the build container by tools/make_fixtures.py: the synthetic reference and reads of a BASELINE
config are generated from a seed, indexed and searched by the reference build (oracle/_ref/lordfast --index, then
oracle/_ref/lordfast_chaindump, the stock search loop with the reference's `alignChain` hook pointed at a recorder), and
every chain handed to alignChain_edlib is stored with the numeric fields, lengths and CRC-32 of the CIGAR / MD of every
Sam_t it produced.  The fixture (fixtures/<name>.npz, git-ignored, travels to the GPU box) holds only chains and
expected records; reference and reads are regenerated from the seed here and checked against the stored CRCs.
"""
from __future__ import annotations

import json
import os
import zlib

import numpy as np

from . import api, sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXDIR = os.path.join(ROOT, "fixtures")                 # the full-size ones: git-ignored, written by build(), travel to the GPU box
GOLDEN = os.path.join(ROOT, "tests", "golden")          # the small ones ("mini*") are committed

# BASELINE.json configs[1..3]; reads are a sample of the configured count where the chain dump would not travel
DATASETS = {
    # configs[1]: 4.6 Mbp (E. coli-sized), 20k x 10 kbp at 12-15 %: all of it
    "config2": dict(ref_len=4_600_000, n_reads=20_000, read_len=10_000, err=(0.12, 0.15), seed=100, sv_frac=0.10, dups=0, num_map=10,
                    what="configs[1]: 4.6 Mbp reference, 20 000 x 10 kbp reads at 12-15 % error, 10 % SV mix"),
    # configs[2]: 64 Mbp (chr20-sized), 15 kbp reads at 15 %: 10k of the 100k reads
    "config3": dict(ref_len=64_000_000, n_reads=10_000, read_len=15_000, err=(0.15, 0.15), seed=300, sv_frac=0.10, dups=0, num_map=10,
                    what="configs[2]: 64 Mbp reference, 15 kbp reads at 15 % error (10 000 of the 100 000 reads), 10 % SV mix"),
    # configs[3] shape: 20 kbp reads, --numMap 10, duplicated segments so that several windows tie (fine mode, several
    # chains per read, wrong-candidate chains with multi-kbp gaps).  A 256 Mbp index is what this container builds in
    # minutes (3.1 Gbp takes hours); t_shift places it at 2.8 Gbp inside a 3.1 Gbp 2-bit reference so that 32-bit
    # reference offsets above 2^31 are exercised.
    "config4": dict(ref_len=256_000_000, n_reads=5_000, read_len=20_000, err=(0.15, 0.15), seed=400, sv_frac=0.10, dups=1200, num_map=10,
                    what="configs[3] shape: 256 Mbp reference with 1 200 duplicated 20-50 kbp segments, 20 kbp reads at 15 % error (5 000 reads), --numMap 10"),
    # small versions for the CPU suite and quick GPU checks
    "mini3": dict(ref_len=2_000_000, n_reads=60, read_len=15_000, err=(0.15, 0.15), seed=301, sv_frac=0.15, dups=0, num_map=10,
                  what="config-3 shape, 60 reads"),
    "mini4": dict(ref_len=2_000_000, n_reads=60, read_len=20_000, err=(0.15, 0.15), seed=401, sv_frac=0.15, dups=20, num_map=10,
                  what="config-4 shape (duplicated segments, --numMap 10), 60 reads"),
}
T_SHIFT_CONFIG4 = 2_800_000_000   # where the 256 Mbp reference sits inside the 3.1 Gbp one
L_PAC_HUMAN = 3_100_000_000


def path(name: str) -> str:
    return os.path.join(GOLDEN if name.startswith("mini") else FIXDIR, "chains_" + name + ".npz" if name.startswith("mini") else name + ".npz")


def available(name: str) -> bool:
    return os.path.exists(path(name))


def make_inputs(name: str):
    """reference + reads of a dataset, from its seed (deterministic)."""
    p = DATASETS[name]
    ref = sim.make_reference_dups(p["ref_len"], p["seed"], p["dups"])
    w = sim.make_workload(p["ref_len"], p["n_reads"], p["read_len"], p["err"][0], p["err"][1], seed=p["seed"], sv_frac=p["sv_frac"], ref=ref)
    return ref, w


class ChainFixture:
    """Stage inputs (reads, reference, chains) + the records the reference produced for them."""

    def __init__(self, name, ref, w, z):
        self.name, self.ref, self.w = name, ref, w
        self.params = json.loads(bytes(z["params"]).decode())
        self.pac = w.pac
        self.reads, self.read_off = w.reads, w.read_off.astype(np.uint64)
        self.contig_off, self.contig_len = np.array([0], dtype=np.int64), np.array([len(ref)], dtype=np.int32)
        self.seeds = np.zeros(len(z["seeds"]), dtype=api.SEED)
        self.seeds["tPos"], self.seeds["qPos"], self.seeds["len"] = z["seeds"][:, 0], z["seeds"][:, 1], z["seeds"][:, 2]
        n = len(z["chain_read"])
        self.chains = np.zeros(n, dtype=api.CHAIN)
        self.chains["seed_off"] = z["chain_seed_off"][:-1]
        self.chains["n_seeds"] = np.diff(z["chain_seed_off"])
        self.chains["read_id"] = z["chain_read"]
        self.chains["is_rev"] = z["chain_rev"]
        self.rec = {k: z["rec_" + k] for k in ("chain", "flag", "pos", "posEnd", "qStart", "qEnd", "nm", "cigar_len", "cigar_crc", "md_len", "md_crc")}
        self.full = json.loads(bytes(z["full"]).decode())   # chain index -> list of {cigar, md} for the first reads

    @property
    def n_reads(self):
        return self.w.n_reads

    @property
    def total_bases(self):
        return self.w.total_bases

    def subset(self, read_lo: int, read_hi: int):
        """chains of the reads [read_lo, read_hi) re-based to a chunk of their own (read sharding)"""
        m = np.flatnonzero((self.chains["read_id"] >= read_lo) & (self.chains["read_id"] < read_hi))
        ch = self.chains[m].copy()
        s_lo = int(ch["seed_off"][0]) if len(ch) else 0
        s_hi = int(ch["seed_off"][-1] + ch["n_seeds"][-1]) if len(ch) else 0
        ch["seed_off"] -= s_lo
        ch["read_id"] -= read_lo
        off = self.read_off[read_lo:read_hi + 1]
        return dict(chains=ch, seeds=self.seeds[s_lo:s_hi], read_off=np.ascontiguousarray(off - off[0]),
                    reads=self.reads[int(off[0]):int(off[-1])], chain_ids=m)

    def compare(self, recs: np.ndarray, text: bytes, t_shift: int = 0, chain_ids=None):
        """lf_sam_record rows of lf_gpu_align_chains against the reference's records.  Returns a list of mismatch
        descriptions (empty = identical).  chain_ids: the fixture chains the call was given (default: all)."""
        R = self.rec
        sel = np.arange(len(R["chain"])) if chain_ids is None else np.flatnonzero(np.isin(R["chain"], chain_ids))
        remap = None
        if chain_ids is not None:
            remap = np.full(len(self.chains), -1, dtype=np.int64)
            remap[chain_ids] = np.arange(len(chain_ids))
        bad = []
        if len(recs) != len(sel):
            return [f"{len(recs)} records, reference has {len(sel)}"]
        mv = memoryview(text)
        for k, j in enumerate(sel):
            r = recs[k]
            exp_chain = int(R["chain"][j]) if remap is None else int(remap[R["chain"][j]])
            got = (int(r["chain_id"]), int(r["flag"]), int(r["pos"]), int(r["posEnd"]), int(r["qStart"]), int(r["qEnd"]), int(r["nmCount"]), int(r["cigar_len"]), int(r["md_len"]))
            exp = (exp_chain, int(R["flag"][j]), int(R["pos"][j]) + t_shift, int(R["posEnd"][j]) + t_shift, int(R["qStart"][j]), int(R["qEnd"][j]), int(R["nm"][j]),
                   int(R["cigar_len"][j]), int(R["md_len"][j]))
            if got != exp:
                bad.append(f"record {k}: {got} != {exp}")
            else:
                co, mo = int(r["cigar_off"]), int(r["md_off"])
                if zlib.crc32(mv[co:co + got[7]]) != int(R["cigar_crc"][j]) or zlib.crc32(mv[mo:mo + got[8]]) != int(R["md_crc"][j]):
                    bad.append(f"record {k} (chain {exp_chain}): CIGAR / MD text differs from the reference's")
            if len(bad) >= 10:
                break
        return bad


def load(name: str) -> ChainFixture:
    if not available(name):
        raise FileNotFoundError(f"{path(name)} missing: run `python tools/make_fixtures.py {name}` in the build container (needs /root/reference)")
    z = np.load(path(name))
    ref, w = make_inputs(name)
    if zlib.crc32(ref.tobytes()) != int(z["ref_crc"]) or zlib.crc32(w.reads.tobytes()) != int(z["reads_crc"]):
        raise RuntimeError(f"fixture {name}: the regenerated reference / reads differ from the ones the chains were dumped for")
    return ChainFixture(name, ref, w, z)
