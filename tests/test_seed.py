"""FM-index seeding: lf_gpu_seed_batch against getLocs_extend_whole_step (src/BWT.cpp:312-394).

CPU part: the Python restatement (oracle/fm_oracle.py) and the numpy index builder (lordfast_b200/fmindex.py) against
the reference's own code (oracle/_ref/libref_shim.so, when it was built here) and against the golden lists it wrote
(tests/golden/seeds_small.npz); the host pipeline + kernels of lf_seed.inl on the test-only emulator against both.
GPU part (-m gpu): the CUDA library against the golden lists, the oracle, and -- where the shim travelled -- the
reference itself at the program's own parameters (k-mer table of 12, 1000 samples per read)."""
import os
import sys
import tempfile

import numpy as np
import pytest

from lordfast_b200 import api, fmindex, sim

import _common
import _oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import fm_oracle  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "seeds_small.npz")
CODE = np.zeros(256, np.uint8)
for _i, _c in enumerate(b"ACGT"):
    CODE[_c] = _i


def small_case(seed=3, ref_len=24_000, n_reads=10, read_len=900):
    """reference with duplicated segments, noisy reads from both strands, and the awkward reads appended"""
    if ref_len >= 20_000:
        ref = sim.make_reference_dups(ref_len, seed=seed, dups=2)
    else:                                          # too short for 4-9 kbp duplications: two exact 300-base copies instead
        ref = sim.make_reference(ref_len, seed=seed)
        ref[ref_len // 2:ref_len // 2 + 300] = ref[100:400]
        ref[ref_len - 700:ref_len - 400] = sim.revcomp(ref[100:400])
    rng = np.random.default_rng(seed)
    reads = []
    for i in range(n_reads):
        L = int(rng.integers(read_len // 2, read_len))
        a = int(rng.integers(0, ref_len - L))
        r = ref[a:a + L].copy()
        hit = rng.random(L) < rng.uniform(0.05, 0.15)
        r[hit] = sim.ACGT[rng.integers(0, 4, size=int(hit.sum()))]
        if i & 1:
            r = sim.revcomp(r)
        reads.append(r)
    at = lambda x: x * ref_len // 24_000         # the positions below are those of the golden case (24 kbp), scaled
    exact = ref[at(5000):at(5000) + 600].copy()                  # a 600-base exact copy: one long match, then containment drops the rest
    reads.append(exact)
    withn = ref[at(9000):at(9000) + 500].copy(); withn[100] = ord("N"); withn[101] = ord("n"); withn[300:303] = ord("-")
    reads.append(withn)
    reads.append(np.char.lower(ref[at(12000):at(12000) + 300].view("S1")).view(np.uint8).copy())   # lower case matches too (nst_nt4_table)
    reads.append(ref[100:110].copy())              # shorter than MIN_ANCHOR_LEN
    reads.append(ref[200:214].copy())              # exactly MIN_ANCHOR_LEN
    reads.append(np.tile(np.frombuffer(b"ACACACACAC", np.uint8), 12))   # low complexity: not in a random reference, or many hits
    reads.append(ref[ref_len - 400:].copy())       # runs into the forward / reverse-complement junction
    reads.append(sim.revcomp(ref[:350]))
    off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
    return ref, np.concatenate(reads).astype(np.uint8), off


PARAM_SETS = [dict(sampling_count=60, min_anchor_len=14, max_ref_hits=1000),
              dict(sampling_count=37, min_anchor_len=12, max_ref_hits=3),      # few hits allowed: repeats are dropped
              dict(sampling_count=1000, min_anchor_len=16, max_ref_hits=50)]  # more samples than bases in the short reads


def lists_equal(a, b):
    for x, y in zip(a, b):
        if not np.array_equal(np.asarray(x), np.asarray(y)):
            return False
    return True


def as_tuple_lists(seeds, off):
    return [[tuple(int(v) for v in s) for s in seeds[int(off[i]):int(off[i + 1])]] for i in range(len(off) - 1)]


needs_ref = pytest.mark.skipif(not (_oracle.have_ref() and hasattr(_oracle.ref(), "ref_fm_load")), reason="reference shim with seeding exports not built (no /root/reference)")


@needs_ref
def test_index_builder_writes_bwa_arrays():
    ref, _, _ = small_case()
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "r.fa")
        _oracle.write_fasta(fa, ref)
        theirs = _oracle.ref_fm_load(fa, 8)
    mine = fmindex.build(CODE[ref], sa_intv=theirs.sa_intv, k_cache=8)
    assert mine.primary == theirs.primary and mine.seq_len == theirs.seq_len and mine.l_pac == theirs.l_pac
    assert np.array_equal(mine.L2, theirs.L2) and np.array_equal(mine.bwt, theirs.bwt) and np.array_equal(mine.sa, theirs.sa)


@needs_ref
def test_oracle_against_reference_seeding():
    ref, reads, off = small_case()
    idx = fm_oracle.TextIndex(CODE[ref])
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "r.fa")
        _oracle.write_fasta(fa, ref)
        _oracle.ref_fm_load(fa, 8)
        rb = reads.tobytes()
        n = 0
        for prm in PARAM_SETS:
            for i in range(len(off) - 1):
                q = rb[int(off[i]):int(off[i + 1])]
                f, r = _oracle.ref_fm_seed(q, **prm)
                of, orr = fm_oracle.seed_read(idx, q, **prm)
                assert [tuple(int(v) for v in s) for s in f] == of, (prm, i)
                assert [tuple(int(v) for v in s) for s in r] == orr, (prm, i)
                n += len(of) + len(orr)
    assert n > 300


def test_oracle_against_golden():
    g = np.load(GOLDEN)
    idx = fm_oracle.TextIndex(CODE[g["ref"]])
    for k, prm in enumerate(PARAM_SETS):
        got = fm_oracle.seed_batch(idx, g["reads"], g["off"], **prm)
        assert lists_equal(got, [g[f"p{k}_{n}"] for n in ("fwd", "fwd_off", "rev", "rev_off")]), prm


def run_lib(g, ref, reads, off, prm, fm=None, resident=False):
    if fm is None:
        fm = fmindex.build(CODE[ref], k_cache=8)
    g.seed_init(fm)
    return g.seed_batch(None if resident else reads, off, **prm)


def test_emu_seeding_matches_golden_and_oracle():
    emu = _common.build_emu()
    gold = np.load(GOLDEN)
    ref, reads, off = gold["ref"], gold["reads"], gold["off"]
    g = api.LfGpu(sim.pack_pac(ref), len(ref), lib_path=emu)
    for k, prm in enumerate(PARAM_SETS):
        got = run_lib(g, ref, reads, off, prm)
        assert lists_equal(got, [gold[f"p{k}_{n}"] for n in ("fwd", "fwd_off", "rev", "rev_off")]), prm
    # a second reference, against the oracle only; the resident-reads form; an empty batch
    ref2, reads2, off2 = small_case(seed=11, ref_len=9000, n_reads=4, read_len=500)
    g2 = api.LfGpu(sim.pack_pac(ref2), len(ref2), lib_path=emu)
    idx = fm_oracle.TextIndex(CODE[ref2])
    prm = dict(sampling_count=50, min_anchor_len=13, max_ref_hits=20)
    want = fm_oracle.seed_batch(idx, reads2, off2, **prm)
    assert lists_equal(run_lib(g2, ref2, reads2, off2, prm), want)
    assert lists_equal(g2.seed_batch(None, off2, **prm), want)
    e = g2.seed_batch(np.zeros(0, np.uint8), np.zeros(1, np.uint64), **prm)
    assert len(e[0]) == 0 and len(e[2]) == 0 and list(e[1]) == [0] and list(e[3]) == [0]


def test_emu_kmer_table_is_the_reference_table():
    """lf_gpu_seed_init without a table derives it on the device; with the reference's table it must give the same seeds"""
    emu = _common.build_emu()
    ref, reads, off = small_case(seed=5, ref_len=6000, n_reads=3, read_len=400)
    fm = fmindex.build(CODE[ref], k_cache=6)
    g = api.LfGpu(sim.pack_pac(ref), len(ref), lib_path=emu)
    g.seed_init(fm)
    table = g.seed_cache()
    # from the definition: the rows of the suffix array that start with the k-mer (beg > end where there are none)
    idx = fm_oracle.TextIndex(CODE[ref])
    for kmer in (0, 1, 4 ** 6 - 1, 1234, 2748):
        digits = [(kmer >> (2 * d)) & 3 for d in range(5, -1, -1)]      # first search character (= last base) first
        pat = bytes(c + 1 for c in reversed(digits))
        k, l = idx.interval(pat)
        if l > k:
            assert (int(table[kmer, 0]), int(table[kmer, 1])) == (k, l - 1), kmer
        else:
            assert table[kmer, 0] > table[kmer, 1], kmer
    if _oracle.have_ref() and hasattr(_oracle.ref(), "ref_fm_load"):
        with tempfile.TemporaryDirectory() as d:
            fa = os.path.join(d, "r.fa")
            _oracle.write_fasta(fa, ref)
            theirs = _oracle.ref_fm_load(fa, 6)
        assert np.array_equal(table, theirs.cache)
        prm = dict(sampling_count=40, min_anchor_len=14, max_ref_hits=100)
        a = g.seed_batch(reads, off, **prm)
        g.seed_init(theirs)
        assert lists_equal(g.seed_batch(reads, off, **prm), a)


def test_seed_argument_checks():
    emu = _common.build_emu()
    ref, reads, off = small_case(seed=5, ref_len=3000, n_reads=1, read_len=200)
    g = api.LfGpu(sim.pack_pac(ref), len(ref), lib_path=emu)
    with pytest.raises(api.LfGpuError):
        g.seed_batch(reads, off)                                   # no index yet
    fm = fmindex.build(CODE[ref], k_cache=6)
    g.seed_init(fm)
    with pytest.raises(api.LfGpuError):
        g.seed_batch(reads, off, min_anchor_len=5)                 # shorter than the k-mer table's k
    with pytest.raises(api.LfGpuError):
        g.seed_batch(reads, off, sampling_count=0)
    bad = fmindex.build(CODE[ref], k_cache=6); bad.sa_intv = 24
    with pytest.raises(api.LfGpuError):
        g.seed_init(bad)


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_seeding_matches_golden_and_oracle():
    gold = np.load(GOLDEN)
    ref, reads, off = gold["ref"], gold["reads"], gold["off"]
    g = api.LfGpu(sim.pack_pac(ref), len(ref))
    for kc in (8, 12):                              # 12: the reference's table size, derived on the device
        fm = fmindex.build(CODE[ref], k_cache=kc)
        for k, prm in enumerate(PARAM_SETS):
            if prm["min_anchor_len"] < kc:
                continue
            got = run_lib(g, ref, reads, off, prm, fm=fm)
            assert lists_equal(got, [gold[f"p{k}_{n}"] for n in ("fwd", "fwd_off", "rev", "rev_off")]), (kc, prm)
    ref2, reads2, off2 = small_case(seed=23, ref_len=40_000, n_reads=40, read_len=1500)
    g2 = api.LfGpu(sim.pack_pac(ref2), len(ref2))
    idx = fm_oracle.TextIndex(CODE[ref2])
    prm = dict(sampling_count=200, min_anchor_len=14, max_ref_hits=1000)
    assert lists_equal(run_lib(g2, ref2, reads2, off2, prm, fm=fmindex.build(CODE[ref2], k_cache=12)), fm_oracle.seed_batch(idx, reads2, off2, **prm))


@pytest.mark.gpu
def test_gpu_seeding_program_parameters_against_reference():
    """lordFAST's own parameters (k-mer table of 12, 1000 samples, 14 / 1000) on 10 kbp reads of a 1 Mbp reference with
    duplications; every list against the reference's own getLocs_extend_whole_step when its build travelled, and
    size-independent properties of the lists otherwise."""
    ref = sim.make_reference_dups(1_000_000, seed=9, dups=4)
    rng = np.random.default_rng(2)
    reads = []
    for i in range(200):
        a = int(rng.integers(0, len(ref) - 10_000))
        r = ref[a:a + 10_000].copy()
        hit = rng.random(10_000) < 0.13
        r[hit] = sim.ACGT[rng.integers(0, 4, size=int(hit.sum()))]
        reads.append(sim.revcomp(r) if i & 1 else r)
    off = (np.arange(201) * 10_000).astype(np.uint64)
    reads = np.concatenate(reads)
    g = api.LfGpu(sim.pack_pac(ref), len(ref))
    codes = CODE[ref]
    fm = fmindex.build(codes, k_cache=12)
    g.seed_init(fm)
    fwd, fo, rev, ro = g.seed_batch(reads, off)
    t = fmindex.both_strands(codes)
    rc = CODE[reads]
    assert len(fwd) + len(rev) > 20_000
    for lst, lo, strand in ((fwd, fo, 0), (rev, ro, 1)):
        for i in (0, 1, 77, 198, 199):
            q = rc[int(off[i]):int(off[i + 1])]
            if strand:
                q = (3 - q[::-1]).astype(np.uint8)       # the reverse list is in coordinates of the reverse-complemented read
            for s in lst[int(lo[i]):int(lo[i + 1])]:
                tp, qp, m = int(s["tPos"]), int(s["qPos"]), int(s["len"])
                assert m >= 14 and np.array_equal(t[tp:tp + m], q[qp:qp + m])   # every seed is an exact match of the stated length
    if _oracle.have_ref() and hasattr(_oracle.ref(), "ref_fm_load"):
        with tempfile.TemporaryDirectory() as d:
            fa = os.path.join(d, "r.fa")
            _oracle.write_fasta(fa, ref)
            theirs = _oracle.ref_fm_load(fa, 12)
            assert np.array_equal(theirs.bwt, fm.bwt) and np.array_equal(theirs.sa, fm.sa)
            assert np.array_equal(g.seed_cache(), theirs.cache)
            rb = reads.tobytes()
            for i in range(200):
                f, r = _oracle.ref_fm_seed(rb[int(off[i]):int(off[i + 1])])
                assert np.array_equal(f, fwd[int(fo[i]):int(fo[i + 1])]), i
                assert np.array_equal(r, rev[int(ro[i]):int(ro[i + 1])]), i


@pytest.mark.gpu
def test_gpu_seeding_config2_chunk_every_read_against_reference():
    """The whole config-2 chunk (20 000 reads of ~10 kbp, 20 M samples, 3.2 M seeds): per read and strand the number of
    seeds and an order-sensitive digest of the list against the reference's own getLocs_extend_whole_step."""
    from lordfast_b200 import fixtures
    if not fixtures.available("config2"):
        pytest.skip("fixtures/config2.npz not present")
    if not (_oracle.have_ref() and hasattr(_oracle.ref(), "ref_fm_seed_batch_digest")):
        pytest.skip("reference shim with the seeding exports not on this box")
    fx = fixtures.load("config2")
    ref, reads, off = fx.ref, np.ascontiguousarray(fx.reads, dtype=np.uint8), fx.read_off.astype(np.uint64)
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "r.fa")
        _oracle.write_fasta(fa, ref)
        theirs = _oracle.ref_fm_load(fa, 12)
    g = api.LfGpu(fx.pac, len(ref))
    theirs.cache = None                      # the library derives the k-mer table; the arrays are bwa's own
    g.seed_init(theirs)
    fwd, fo, rev, ro = g.seed_batch(reads, off)
    dig, cnt, _ = _oracle.ref_fm_seed_digests(reads, off)
    assert np.array_equal(np.diff(fo.astype(np.int64)), cnt[:, 0]) and np.array_equal(np.diff(ro.astype(np.int64)), cnt[:, 1])
    assert np.array_equal(_oracle.seed_list_digests(fwd, fo), dig[:, 0]) and np.array_equal(_oracle.seed_list_digests(rev, ro), dig[:, 1])
    assert len(fwd) + len(rev) > 2_000_000


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_emu_seeding_random_cases_against_oracle(seed):
    """Random small references (with repeats and a palindromic stretch), reads with substitutions, indels, N runs and
    lower case, random parameters at the edges (one sample, min_anchor_len = k of the table, a single allowed hit)."""
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(1500, 5000))
    ref = sim.make_reference(n, seed=50 + seed)
    a = int(rng.integers(0, n - 400)); b = int(rng.integers(0, n - 400))
    ref[b:b + 200] = ref[a:a + 200]                                  # an exact repeat
    c = int(rng.integers(0, n - 100)); ref[c + 40:c + 80] = sim.revcomp(ref[c:c + 40])   # a reverse-complement palindrome
    reads = []
    for i in range(8):
        L = int(rng.integers(1, 400))
        s = int(rng.integers(0, n - L))
        r = ref[s:s + L].copy()
        hit = rng.random(L) < rng.uniform(0.0, 0.2)
        r[hit] = sim.ACGT[rng.integers(0, 4, size=int(hit.sum()))]
        if L > 30 and i % 3 == 0:
            p = int(rng.integers(0, L - 10)); r = np.concatenate([r[:p], r[p + int(rng.integers(1, 6)):]])        # deletion
        if L > 30 and i % 3 == 1:
            p = int(rng.integers(0, L - 10)); r = np.concatenate([r[:p], sim.ACGT[rng.integers(0, 4, size=3)], r[p:]])   # insertion
        if len(r) > 20 and i % 4 == 2:
            p = int(rng.integers(0, len(r) - 5)); r[p:p + int(rng.integers(1, 5))] = ord("N")
        if i % 5 == 4:
            r = np.char.lower(r.view("S1")).view(np.uint8).copy()
        if i & 1:
            r = sim.revcomp(np.where(np.isin(r, sim.ACGT), r, ord("A")).astype(np.uint8)) if not np.isin(r, sim.ACGT).all() else sim.revcomp(r)
        reads.append(r.astype(np.uint8))
    reads.append(np.zeros(0, np.uint8))                               # an empty read in the middle of the batch
    reads.append(np.full(40, ord("N"), np.uint8))
    order = rng.permutation(len(reads))
    reads = [reads[i] for i in order]
    off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
    allr = np.concatenate(reads) if sum(len(r) for r in reads) else np.zeros(0, np.uint8)
    kc = int(rng.integers(4, 8))
    g = api.LfGpu(sim.pack_pac(ref), len(ref), lib_path=_common.build_emu())
    g.seed_init(fmindex.build(CODE[ref], sa_intv=int(rng.choice([1, 4, 32])), k_cache=kc))
    idx = fm_oracle.TextIndex(CODE[ref])
    have_shim = _oracle.have_ref() and hasattr(_oracle.ref(), "ref_fm_load")
    if have_shim:
        with tempfile.TemporaryDirectory() as d:
            fa = os.path.join(d, "r.fa")
            _oracle.write_fasta(fa, ref)
            _oracle.ref_fm_load(fa, kc)
    for prm in (dict(sampling_count=1, min_anchor_len=kc, max_ref_hits=1000),
                dict(sampling_count=int(rng.integers(2, 90)), min_anchor_len=int(rng.integers(kc, 20)), max_ref_hits=int(rng.integers(1, 6))),
                dict(sampling_count=500, min_anchor_len=14, max_ref_hits=1000)):
        want = fm_oracle.seed_batch(idx, allr, off, **prm)
        assert lists_equal(g.seed_batch(allr, off, **prm), want), prm
        if have_shim:                                                 # ... and the reference itself on the same reads
            rb = allr.tobytes()
            for i in range(len(off) - 1):
                f, r = _oracle.ref_fm_seed(rb[int(off[i]):int(off[i + 1])], **prm)
                assert np.array_equal(f, want[0][int(want[1][i]):int(want[1][i + 1])]) and np.array_equal(r, want[2][int(want[3][i]):int(want[3][i + 1])]), (prm, i)
    g.close()
