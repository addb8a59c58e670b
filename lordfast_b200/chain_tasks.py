"""Round-1 task derivation from chains (src/LordFAST.cpp:1820-1833, :1902-1941,
:2157-2168 of the reference): every head / gap / tail alignment a chain needs is known from the chain
alone, so a whole chunk's tasks are emitted in one vectorised pass.

Used by bench.py and the tests to build lf_align_task arrays; the C++ chain driver
(lordfast_b200/csrc/lf_chain.inl) does the same per chain and also handles the result-dependent
rounds 2 and 3.
"""
from __future__ import annotations

import numpy as np

from .api import ALIGN_TASK, LF_F_NO_PATH, LF_F_READ_REV, LF_F_REVERSE_BOTH, LF_MODE_NW, LF_MODE_SHW

CLIP_LEN = 500  # _pf_clipLen (src/LordFAST.cpp:88)

KIND_HEAD, KIND_GAP, KIND_TAIL = 0, 1, 2


def round1_tasks(seeds: np.ndarray, seed_off: np.ndarray, is_rev: np.ndarray, read_len: np.ndarray,
                 contig_off: np.ndarray, contig_len: np.ndarray, read_id: np.ndarray | None = None, long_ends_nopath: bool = True):
    """seeds uint32[m,3] (tPos,qPos,len); one chain per entry of seed_off[:-1].
    Returns (tasks, chain_of_task, kind_of_task) with tasks in chain order: head, gaps, tail."""
    n = len(seed_off) - 1
    if read_id is None:
        read_id = np.arange(n, dtype=np.uint32)
    s = seeds.astype(np.int64)
    first = seed_off[:-1].astype(np.int64)
    last = seed_off[1:].astype(np.int64) - 1
    rl = read_len.astype(np.int64)
    strand = np.where(is_rev.astype(bool), LF_F_READ_REV, 0).astype(np.uint16)
    # contig of a chain: bwt_get_chr_boundaries on the midpoint of first/last seed (src/BWT.cpp:653-660)
    mid = (s[first, 0] + s[last, 0]) >> 1
    rid = np.clip(np.searchsorted(contig_off, mid, side="right") - 1, 0, len(contig_off) - 1)
    chr_beg = contig_off[rid].astype(np.int64)
    chr_end = chr_beg + contig_len[rid].astype(np.int64) - 1

    # gaps between consecutive seeds of the same chain
    nseed = len(s)
    chain_of_seed = np.repeat(np.arange(n), (last - first + 1))
    has_next = np.ones(nseed, dtype=bool)
    has_next[last] = False
    i0 = np.flatnonzero(has_next)
    qs = s[i0, 1] + s[i0, 2]
    ts = s[i0, 0] + s[i0, 2]
    ql = s[i0 + 1, 1] - qs
    tl = s[i0 + 1, 0] - ts
    g = (ql > 0) & (tl > 0)
    gap = np.zeros(int(g.sum()), dtype=ALIGN_TASK)
    gc = chain_of_seed[i0][g]
    gap["read_id"], gap["q_off"], gap["q_len"], gap["t_off"], gap["t_len"] = read_id[gc], qs[g], ql[g], ts[g], tl[g]
    gap["flags"], gap["mode"] = strand[gc], LF_MODE_NW
    gap_pos = i0[g]  # seed index: orders the gap inside its chain

    # head: SHW of the reversed read head against the reversed 20-longer reference slice
    a = s[first, 1]
    hm = (a > 0) & (s[first, 0] - (a + 20) >= chr_beg)
    hc = np.flatnonzero(hm)
    head = np.zeros(len(hc), dtype=ALIGN_TASK)
    head["read_id"], head["q_off"], head["q_len"] = read_id[hc], 0, a[hc]
    head["t_off"], head["t_len"] = s[first[hc], 0] - (a[hc] + 20), a[hc] + 20
    # heads / tails longer than _pf_clipLen: distance and end only in round 1, as lf_gpu_align_chains issues them (their
    # path is thrown away whenever the clip test fires and the extension shortens them; otherwise round 3 computes it)
    head["flags"], head["mode"] = strand[hc] | LF_F_REVERSE_BOTH | np.where((a[hc] > CLIP_LEN) & long_ends_nopath, LF_F_NO_PATH, 0).astype(np.uint16), LF_MODE_SHW

    # tail
    qs_t = s[last, 1] + s[last, 2]
    b = rl - qs_t
    ts_t = s[last, 0] + s[last, 2]
    tm = (b > 0) & (ts_t + (b + 20) - 1 <= chr_end)
    tc = np.flatnonzero(tm)
    tail = np.zeros(len(tc), dtype=ALIGN_TASK)
    tail["read_id"], tail["q_off"], tail["q_len"] = read_id[tc], qs_t[tc], b[tc]
    tail["t_off"], tail["t_len"] = ts_t[tc], b[tc] + 20
    tail["flags"], tail["mode"] = strand[tc] | np.where((b[tc] > CLIP_LEN) & long_ends_nopath, LF_F_NO_PATH, 0).astype(np.uint16), LF_MODE_SHW

    tasks = np.concatenate([head, gap, tail])
    chain = np.concatenate([hc, gc, tc])
    kind = np.concatenate([np.full(len(hc), KIND_HEAD), np.full(len(gc), KIND_GAP), np.full(len(tc), KIND_TAIL)])
    pos = np.concatenate([first[hc] - 1, gap_pos, last[tc] + 1])  # position along the chain
    order = np.lexsort((pos, chain))
    return np.ascontiguousarray(tasks[order]), chain[order], kind[order]


def workload_tasks(w, long_ends_nopath: bool = True):
    """Round-1 tasks of a sim.Workload (one chain per read); long_ends_nopath=False asks for the path of every task."""
    read_len = np.diff(w.read_off)
    return round1_tasks(w.seeds, w.seed_off, w.is_rev, read_len, w.contig_off, w.contig_len, long_ends_nopath=long_ends_nopath)
