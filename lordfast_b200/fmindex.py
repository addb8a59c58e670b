"""bwa's FM index of a (synthetic) reference as numpy arrays, in the memory layout lordFAST loads it in.

lordFAST seeds on `bwa index` files (bwt_t, lib/bwa/bwt.h:44-57, built by lib/bwa/bwtindex.c:bwa_index and
bwt.c:bwt_cal_sa) plus its own 12-mer table (src/BWT.cpp:60-138).  The GPU seeding stage (lf_gpu_seed_init) takes those
arrays as they are.  This module builds the same arrays for a reference that only exists as a numpy array -- tests,
smoke() and benches have no `bwa index` on the GPU box -- by the textbook route (suffix array by prefix doubling),
not bwa's (BWT-SW / IS): the result is a function of the text alone, and tests/test_seed_oracle.py compares it word
for word with what the reference's own code writes.

Layout (all of it bwa's):
  text     T = forward strand + reverse complement, 2 * l_pac bases (bntseq.c: for_only = 0)
  bwt      BWT of T$ with the $ removed (`primary` = its row); per 128 bases four 64-bit running counts, then eight
           32-bit words of 16 bases, first base in the top two bits; one more count record after the last block
           (bwt.c:bwt_bwtupdate_core)
  L2       cumulative base counts
  sa       SA[i] for every sa_intv-th row i, 64-bit, sa[0] = -1 (bwt_cal_sa)
"""
from __future__ import annotations

import dataclasses

import numpy as np

OCC_INTERVAL = 128


@dataclasses.dataclass
class FmIndex:
    bwt: np.ndarray        # uint32
    sa: np.ndarray         # uint64
    primary: int
    L2: np.ndarray         # uint64[5]
    seq_len: int
    sa_intv: int
    l_pac: int
    k_cache: int = 12
    cache: np.ndarray | None = None   # uint64[4^k_cache, 2] (beg, end) or None: built on the device


def both_strands(fwd: np.ndarray) -> np.ndarray:
    """fwd: base codes 0..3.  bwa appends the reverse complement (bntseq.c:bns_fasta2bntseq)."""
    fwd = np.asarray(fwd, dtype=np.uint8)
    return np.concatenate([fwd, (3 - fwd[::-1]).astype(np.uint8)])


def suffix_array(t: np.ndarray) -> np.ndarray:
    """Suffix array of t + $ ($ smallest): n + 1 rows, row 0 is the $ suffix.  Prefix doubling."""
    n = len(t)
    rank = np.concatenate([t.astype(np.int64) + 1, np.zeros(1, np.int64)])
    sa = np.argsort(rank, kind="stable")
    k = 1
    while True:
        r2 = np.zeros(n + 1, np.int64)
        if k <= n:
            r2[: n + 1 - k] = rank[k:]
        key = rank * (n + 2) + r2
        sa = np.argsort(key, kind="stable")
        ks = key[sa]
        newrank = np.empty(n + 1, np.int64)
        newrank[sa] = np.concatenate([[0], np.cumsum(ks[1:] != ks[:-1])])
        rank = newrank
        if rank.max() == n:
            return sa.astype(np.int64)
        k *= 2


def build(fwd: np.ndarray, sa_intv: int = 32, k_cache: int = 12) -> FmIndex:
    t = both_strands(fwd)
    n = len(t)
    sa = suffix_array(t)
    primary = int(np.nonzero(sa == 0)[0][0])
    keep = sa != 0
    b0 = t[sa[keep] - 1]                                    # the $-removed BWT string, n bases
    counts = np.bincount(t, minlength=4).astype(np.uint64)
    L2 = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    # 16 bases per word, first base on top
    pad = (-n) % 16
    bp = np.concatenate([b0, np.zeros(pad, np.uint8)]).reshape(-1, 16).astype(np.uint32)
    words = np.zeros(len(bp), np.uint32)
    for j in range(16):
        words |= bp[:, j] << np.uint32(30 - 2 * j)
    # running counts at every 128th base, and once more at the end
    n_blocks = (n + OCC_INTERVAL - 1) // OCC_INTERVAL
    onehot = (b0[:, None] == np.arange(4, dtype=np.uint8)[None, :])
    cum = np.concatenate([np.zeros((1, 4), np.uint64), np.cumsum(onehot, axis=0, dtype=np.uint64)])
    out = np.zeros(n_blocks * 16 + 8, np.uint32)
    blk = out[: n_blocks * 16].reshape(n_blocks, 16)
    blk[:, :8] = cum[np.arange(n_blocks) * OCC_INTERVAL].view(np.uint32).reshape(n_blocks, 8)
    wpad = np.concatenate([words, np.zeros(n_blocks * 8 - len(words), np.uint32)])
    blk[:, 8:] = wpad.reshape(n_blocks, 8)
    out[n_blocks * 16:] = cum[n].view(np.uint32)
    # bwa keeps only the words that exist: (n + 15) / 16 data words + (n_blocks + 1) count records
    size = (n + 15) // 16 + (n_blocks + 1) * 8
    last_words = (n + 15) // 16 - (n_blocks - 1) * 8        # data words of the last block
    if last_words < 8:                                      # the end record follows the last data word directly
        tail = out[n_blocks * 16:].copy()
        cut = (n_blocks - 1) * 16 + 8 + last_words
        out[cut: cut + 8] = tail
    bwt = out[:size].copy()
    n_sa = (n + sa_intv) // sa_intv
    sas = np.zeros(n_sa, np.uint64)
    rows = np.arange(0, n + 1, sa_intv)
    sas[rows // sa_intv] = sa[rows].astype(np.uint64)
    sas[0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return FmIndex(bwt=bwt, sa=sas, primary=primary, L2=L2, seq_len=n, sa_intv=sa_intv, l_pac=len(fwd), k_cache=k_cache)
