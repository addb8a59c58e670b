"""CPU restatement of lordFAST's seeding step, getLocs_extend_whole_step (src/BWT.cpp:312-394).

TEST INFRASTRUCTURE ONLY: the checker for lf_gpu_seed_batch, never a fallback for it.  Only tests/, smoke() and the
cpu_baseline leg of the benches import this module.

Written from the definition, without an FM index: what the reference computes through bwa's backward search
(bwt_count_exact_cached, src/BWT.cpp:265-298) is a property of the text alone --
  * the text is T = forward strand + reverse complement (bwa indexes both, lib/bwa/bntseq.c),
  * count(P) = the rows of the suffix array of T$ whose suffixes start with P, a contiguous range,
  * bwt_sa(row) = that suffix's start,
so this module keeps the plain suffix array and finds the ranges by binary search over it.  The control flow around it
follows the reference line by line (sample positions in double precision :320-322 / :389-390, longest match :328-342,
the hit-count and containment tests :345, strand and coordinates :351-371, Seed_t's 20- and 12-bit fields
src/LordFAST.h:30-35).  Pinned against the reference's own code (oracle/_ref/libref_shim.so: ref_fm_seed) by
tests/test_seed.py, and against the golden lists in tests/golden/seeds_*.npz that the same shim wrote.
"""
from __future__ import annotations

import numpy as np

_NT4 = np.full(256, 4, dtype=np.uint8)          # nst_nt4_table (lib/bwa/bntseq.c:47-64), 5 ('-') folded into 4
for _i, _c in enumerate(b"ACGT"):
    _NT4[_c] = _i
    _NT4[_c + 32] = _i


def plain_suffix_array(t: np.ndarray, key_len: int = 384) -> np.ndarray:
    """Rows of the suffix array of t + $ ($ smallest).  Plain sort of suffixes (small texts) -- nothing shared with the
    product's index builder: sort on the first key_len bytes, then order the (rare) ties by whole suffixes."""
    import functools
    n = len(t)
    tb = (t.astype(np.uint8) + 1).tobytes() + b"\x00"
    rows = sorted(range(n + 1), key=lambda p: tb[p:p + key_len])
    i = 0
    while i < n:
        j = i + 1
        ki = tb[rows[i]:rows[i] + key_len]
        while j <= n and tb[rows[j]:rows[j] + key_len] == ki:
            j += 1
        if j - i > 1:
            rows[i:j] = sorted(rows[i:j], key=functools.cmp_to_key(lambda a, b: -1 if tb[a:] < tb[b:] else 1))
        i = j
    return np.array(rows, dtype=np.int64)


class TextIndex:
    def __init__(self, fwd_codes: np.ndarray, sa: np.ndarray | None = None):
        fwd = np.asarray(fwd_codes, dtype=np.uint8)
        self.l_pac = len(fwd)
        self.t = np.concatenate([fwd, (3 - fwd[::-1]).astype(np.uint8)])
        self.tb = (self.t + 1).tobytes()             # bases as bytes 1..4: every real base sorts after the $ (0)
        self.sa = plain_suffix_array(self.t) if sa is None else np.asarray(sa, dtype=np.int64)

    def _bound(self, pat: bytes, lo: int, hi: int, upper: bool) -> int:
        tb, sa, L = self.tb, self.sa, len(pat)
        while lo < hi:
            mid = (lo + hi) >> 1
            p = int(sa[mid])
            s = tb[p:p + L]
            if s < pat or (upper and s == pat):
                lo = mid + 1
            else:
                hi = mid
        return lo

    def interval(self, pat: bytes, lo: int = 0, hi: int | None = None):
        """rows [k, l) whose suffixes start with pat, searched inside [lo, hi)"""
        hi = len(self.sa) if hi is None else hi
        k = self._bound(pat, lo, hi, False)
        return k, self._bound(pat, k, hi, True)


def seed_read(idx: TextIndex, read: bytes, sampling_count: int = 1000, min_anchor_len: int = 14, max_ref_hits: int = 1000):
    """(forward list, reverse list) of (tPos, qPos, len) for one read."""
    qlen = len(read)
    codes = _NT4[np.frombuffer(read, dtype=np.uint8)]
    bad = np.nonzero(codes > 3)[0]
    pat_all = (codes + 1).astype(np.uint8).tobytes()
    step = float(qlen) / sampling_count
    seed_pos, pos, last = 0.0, 0, 0
    fwd, rev = [], []
    for _ in range(sampling_count):
        # the longest L >= min_anchor_len for which read[pos : pos + L] occurs (a non-ACGT base or the end of the read stops it)
        stop = qlen
        j = np.searchsorted(bad, pos)
        if j < len(bad):
            stop = int(bad[j])
        m, k, l = 0, 0, 0
        if pos + min_anchor_len <= stop:
            k, l = idx.interval(pat_all[pos:pos + min_anchor_len])
            if l > k:
                m = min_anchor_len
                while pos + m + 1 <= stop:
                    k2, l2 = idx.interval(pat_all[pos:pos + m + 1], k, l)
                    if l2 <= k2:
                        break
                    k, l, m = k2, l2, m + 1
        occ = l - k if m else 0
        if not m:
            m = min_anchor_len          # the reference's m stays at MIN_ANCHOR_LEN when nothing matches (occ = 0 decides)
        if 0 < occ < max_ref_hits and pos + m > last:
            for row in range(k, l):
                sapos = int(idx.sa[row])
                if sapos >= idx.l_pac:
                    rev.append(((2 * idx.l_pac - sapos - m) & 0xFFFFFFFF, (qlen - pos - m) & 0xFFFFF, m & 0xFFF))
                else:
                    fwd.append((sapos, pos & 0xFFFFF, m & 0xFFF))
            last = pos + m
        seed_pos += step
        pos = int(seed_pos)
    return fwd, rev


def seed_batch(idx: TextIndex, reads: np.ndarray, offsets: np.ndarray, **kw):
    """Same return shape as LfGpu.seed_batch: (fwd, fwd_off, rev, rev_off) with structured (tPos, qPos, len) arrays."""
    dt = np.dtype([("tPos", "<u4"), ("qPos", "<u4"), ("len", "<u4")])
    f_all, r_all, fo, ro = [], [], [0], [0]
    rb = np.asarray(reads, dtype=np.uint8).tobytes()
    for i in range(len(offsets) - 1):
        f, r = seed_read(idx, rb[int(offsets[i]):int(offsets[i + 1])], **kw)
        f_all += f; r_all += r
        fo.append(len(f_all)); ro.append(len(r_all))
    mk = lambda lst: np.array(lst, dtype=dt) if lst else np.zeros(0, dtype=dt)
    return mk(f_all), np.array(fo, dtype=np.uint64), mk(r_all), np.array(ro, dtype=np.uint64)
