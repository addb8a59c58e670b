"""Deterministic synthetic inputs for the alignment stage.

reference : i.i.d. uniform ACGT, one or more contigs
reads     : uniform start, random strand, per-base error channel with total rate e split
            sub:ins:del = 10:60:30 (PacBio-CLR-like), upper-case ACGT, no N
SV mix    : junk head / junk tail (1-2 kbp), 600 bp deletion, 600 bp insertion, 1.5 kbp inversion,
            so that the soft-clip and split triggers (ksw_extend paths) fire
chains    : a front-end model -- every exact-match run >= MIN_ANCHOR_LEN between error events is an
            anchor, thinned to the density lordFAST's own seeding reaches (measured here with
            oracle/_ref/lordfast_chaindump: ~112 anchors per 10 kbp read at 15 % error, gap
            median 46 / mean 68 bp).  Real chains
            from the reference front-end come from oracle/_ref/lordfast_chaindump instead
            (tests/golden/make_golden.py).

Everything is numpy-vectorised so that the 20k x 10 kbp workload is generated in seconds.
"""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = ord("N")
for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[a] = b

MIN_ANCHOR_LEN = 14  # lordFAST default -k (src/CommandLineParser.cpp)
MAX_SEED_LEN = 4095  # Seed_t.len is a 12-bit field (src/LordFAST.h:30-35)


def make_reference(n: int, seed: int = 1) -> np.ndarray:
    """ASCII uint8 array of n uniform ACGT bases."""
    rng = np.random.default_rng(seed)
    return ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def make_reference_dups(n: int, seed: int = 1, dups: int = 0) -> np.ndarray:
    """make_reference plus `dups` segments of 20-50 kbp (4-9 kbp in small references) copied elsewhere at 2-5 %
    divergence, so that several candidate windows tie and lordFAST's fine mode / --numMap path is taken."""
    ref = make_reference(n, seed)
    if dups:
        rng = np.random.default_rng(seed + 31)
        for _ in range(dups):
            L = int(rng.integers(20_000, 50_000)) if n > 400_000 else int(rng.integers(4_000, 9_000))
            a, b = int(rng.integers(0, n - L)), int(rng.integers(0, n - L))
            seg = ref[a:a + L].copy()
            hit = rng.random(L) < rng.uniform(0.02, 0.05)
            seg[hit] = ACGT[(np.searchsorted(ACGT, seg[hit]) + rng.integers(1, 4, size=int(hit.sum()))) % 4]
            ref[b:b + L] = seg
    return ref


def revcomp(a: np.ndarray) -> np.ndarray:
    return _COMP[a[::-1]]


def pack_pac(ref: np.ndarray) -> np.ndarray:
    """bwa .pac layout: base l in byte l>>2 at bits ((~l)&3)<<1 (lib/bwa/bntseq.c:224-225);
    length l/4+1 as `bwa index` writes it."""
    code = np.zeros(256, dtype=np.uint8)
    for i, c in enumerate(b"ACGT"):
        code[c] = i
    c = code[ref]
    n = len(c)
    pad = (-n) % 4
    if pad:
        c = np.concatenate([c, np.zeros(pad, dtype=np.uint8)])
    c = c.reshape(-1, 4)
    out = (c[:, 0] << 6) | (c[:, 1] << 4) | (c[:, 2] << 2) | c[:, 3]
    return np.concatenate([out.astype(np.uint8), np.zeros(n // 4 + 1 - len(out), dtype=np.uint8)])


def _channel(src: np.ndarray, e: float, rng) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Push `src` through the error channel.  Returns (out, qpos_of_src, clean) where
    qpos_of_src[i] is the output index of source base i (or of the next emitted base if i was
    deleted) and clean[i] says base i came through unchanged with no insertion before it."""
    L = len(src)
    e_sub, e_ins, e_del = 0.1 * e, 0.6 * e, 0.3 * e
    u = rng.random(L)
    deleted = u < e_del
    subbed = (u >= e_del) & (u < e_del + e_sub)
    # insertion runs are geometric: P(len >= k) = e_ins' ^ k with mean e_ins per base
    p_open = e_ins / (1.0 + e_ins)
    ins_len = np.where(rng.random(L) < p_open, rng.geometric(1.0 - p_open, size=L), 0)
    emit = (~deleted).astype(np.int64) + ins_len  # inserted bases go BEFORE the base
    starts = np.cumsum(emit) - emit
    total = int(emit.sum())
    out = ACGT[rng.integers(0, 4, size=total, dtype=np.uint8)]
    base_pos = starts + ins_len  # where the (kept) source base lands
    kept = ~deleted
    vals = src.copy()
    if subbed.any():
        # substitute with a different base
        code = np.searchsorted(ACGT, vals[subbed])
        vals[subbed] = ACGT[(code + rng.integers(1, 4, size=int(subbed.sum()))) % 4]
    out[base_pos[kept]] = vals[kept]
    clean = kept & ~subbed & (ins_len == 0)
    return out, base_pos, clean


ANCHOR_KEEP = 0.68  # lordFAST samples 1000 positions per read and finds ~112 of the ~165 runs


def _anchors(qpos: np.ndarray, clean: np.ndarray, t0: int, q0: int, min_len: int, rng=None):
    """Maximal clean runs -> (tPos, qPos, len) anchors.  A run may start on a base that had an
    insertion before it (the insertion precedes the base), so the run's first base only needs to
    be kept and unsubstituted; we use the stricter `clean` for simplicity and drop that base."""
    if len(clean) == 0:
        return np.zeros((0, 3), dtype=np.int64)
    d = np.diff(np.concatenate([[0], clean.astype(np.int8), [0]]))
    s = np.flatnonzero(d == 1)
    e = np.flatnonzero(d == -1)
    ln = e - s
    ok = ln >= min_len
    if rng is not None:
        ok &= rng.random(len(ln)) < ANCHOR_KEEP
    s, ln = s[ok], np.minimum(ln[ok], MAX_SEED_LEN)
    return np.stack([t0 + s, q0 + qpos[s], ln], axis=1)


SV_KINDS = ("junk_head", "junk_tail", "deletion", "insertion", "inversion")


def simulate_read(ref: np.ndarray, read_len: int, e: float, rng, kind: str = "plain", min_anchor: int = MIN_ANCHOR_LEN,
                  lo: int = 0, hi: int | None = None):
    """One read in REFERENCE orientation plus its chain model.  Returns (oriented_read, seeds[n,3])."""
    hi = len(ref) if hi is None else hi
    span = int(read_len / (1.0 + 0.3 * e))  # source bases so that the read comes out ~read_len
    sv = 0
    if kind == "deletion":
        sv = 600
    elif kind in ("inversion", "inversion_del"):
        sv = 1500  # room for either variant
    margin = 64
    start = int(rng.integers(lo + margin, max(lo + margin + 1, hi - span - sv - margin)))
    if kind in ("plain", "junk_head", "junk_tail"):
        out, qpos, clean = _channel(ref[start:start + span], e, rng)
        seeds = _anchors(qpos, clean, start, 0, min_anchor, rng)
        if kind != "plain":
            junk = ACGT[rng.integers(0, 4, size=int(rng.integers(1000, 2000)), dtype=np.uint8)]
            if kind == "junk_head":
                seeds[:, 1] += len(junk)
                out = np.concatenate([junk, out])
            else:
                out = np.concatenate([out, junk])
        return out, seeds
    half = span // 2
    a, qa, ca = _channel(ref[start:start + half], e, rng)
    sa = _anchors(qa, ca, start, 0, min_anchor, rng)
    if kind == "deletion":
        b0 = start + half + 600
        mid = np.zeros(0, dtype=np.uint8)
    elif kind == "insertion":
        b0 = start + half
        mid = ACGT[rng.integers(0, 4, size=600, dtype=np.uint8)]
    elif kind == "inversion_del":
        # 800 bp inverted (nearly error-free) followed by a 240 bp deletion: the gap is dissimilar enough for lordFAST's
        # split test (|q-t| >= 80, similarity < 0.40) and its reverse complement aligns above 0.60, so the
        # accepted-inversion branch (src/LordFAST.cpp:2040-2074) is taken
        b0 = start + half + 800 + 240
        mid, _, _ = _channel(revcomp(ref[start + half:start + half + 800]), 0.02, rng)
    else:  # inversion: 1.5 kbp of the reference comes through reverse-complemented
        b0 = start + half + 1500
        mid, _, _ = _channel(revcomp(ref[start + half:start + half + 1500]), e, rng)
    b, qb, cb = _channel(ref[b0:b0 + (span - half)], e, rng)
    sb = _anchors(qb, cb, b0, len(a) + len(mid), min_anchor, rng)
    return np.concatenate([a, mid, b]), np.concatenate([sa, sb])


class Workload:
    """Stage inputs for a batch of reads: forward reads (as sequenced), one chain per read.

    reads      uint8[total]   concatenated read bytes
    read_off   int64[n+1]
    is_rev     uint8[n]       1: the chain refers to the reverse complement of the stored read
    seeds      uint32[m,3]    (tPos, qPos, len) in oriented-read / forward-reference coordinates
    seed_off   int64[n+1]
    """

    def __init__(self, ref, pac, reads, read_off, is_rev, seeds, seed_off, kinds):
        self.ref, self.pac, self.reads, self.read_off = ref, pac, reads, read_off
        self.is_rev, self.seeds, self.seed_off, self.kinds = is_rev, seeds, seed_off, kinds
        self.contig_off = np.array([0], dtype=np.int64)
        self.contig_len = np.array([len(ref)], dtype=np.int32)

    @property
    def n_reads(self):
        return len(self.is_rev)

    @property
    def total_bases(self):
        return int(self.read_off[-1])

    def oriented(self, i: int) -> np.ndarray:
        r = self.reads[self.read_off[i]:self.read_off[i + 1]]
        return revcomp(r) if self.is_rev[i] else r

    def chain(self, i: int) -> np.ndarray:
        return self.seeds[self.seed_off[i]:self.seed_off[i + 1]]


def make_workload(ref_len: int, n_reads: int, read_len: int, err_lo: float, err_hi: float, seed: int = 1,
                  sv_frac: float = 0.10, ref: np.ndarray | None = None, sv_kinds=SV_KINDS) -> Workload:
    rng = np.random.default_rng(seed + 7919)
    if ref is None:
        ref = make_reference(ref_len, seed)
    reads, seeds, is_rev, kinds = [], [], [], []
    for i in range(n_reads):
        e = float(rng.uniform(err_lo, err_hi))
        kind = "plain"
        if rng.random() < sv_frac:
            kind = sv_kinds[int(rng.integers(0, len(sv_kinds)))]
        for _ in range(20):  # a chain needs at least two anchors
            o, s = simulate_read(ref, read_len, e, rng, kind)
            if len(s) >= 2:
                break
        rev = int(rng.integers(0, 2))
        reads.append(revcomp(o) if rev else o)
        seeds.append(s)
        is_rev.append(rev)
        kinds.append(kind)
    read_off = np.zeros(n_reads + 1, dtype=np.int64)
    read_off[1:] = np.cumsum([len(r) for r in reads])
    seed_off = np.zeros(n_reads + 1, dtype=np.int64)
    seed_off[1:] = np.cumsum([len(s) for s in seeds])
    return Workload(ref, pack_pac(ref), np.concatenate(reads), read_off, np.array(is_rev, dtype=np.uint8),
                    np.concatenate(seeds).astype(np.uint32), seed_off, kinds)


def mutate_pair(target: np.ndarray, d: float, rng) -> np.ndarray:
    """Kernel-microbench query: `target` pushed through the channel at divergence d (config 5)."""
    out, _, _ = _channel(target, d, rng)
    return out if len(out) else target[:1].copy()
