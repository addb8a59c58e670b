"""ctypes access to the checkers: oracle/liblforacle.so (our CPU restatement) and, when it was
built in this container, oracle/_ref/libref_shim.so (the reference's own edlib / ksw /
alignChain_edlib).  Test infrastructure only -- nothing in lordfast_b200/ imports this."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liblforacle.so")
REF_SHIM_SO = os.path.join(ORACLE_DIR, "_ref", "libref_shim.so")

CLIP_MAT = (C.c_int8 * 25)(*[2 if (i == j and i < 4) else (0 if (i == 4 or j == 4) else -16)
                            for i in range(5) for j in range(5)])


class AlignOut(C.Structure):
    _fields_ = [("edit_distance", C.c_int), ("end_location", C.c_int), ("n_ops", C.c_int)]


class Sam(C.Structure):
    _fields_ = [("flag", C.c_uint32), ("pos", C.c_uint32), ("posEnd", C.c_uint32), ("qStart", C.c_uint32),
                ("qEnd", C.c_uint32), ("nmCount", C.c_int32), ("cigar", C.c_void_p), ("md", C.c_void_p)]


class Seed(C.Structure):
    _fields_ = [("tPos", C.c_uint32), ("qPos", C.c_uint32), ("len", C.c_uint32)]


class Ref(C.Structure):
    _fields_ = [("pac", C.c_void_p), ("l_pac", C.c_int64), ("n_contigs", C.c_int),
                ("contig_off", C.POINTER(C.c_int64)), ("contig_len", C.POINTER(C.c_int32))]


class Stats(C.Structure):
    _fields_ = [("n_align", C.c_int64), ("n_extend", C.c_int64), ("cells_align", C.c_int64), ("cells_extend", C.c_int64)]


def build_oracle():
    src = os.path.join(ORACLE_DIR, "lf_oracle.c")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    return ORACLE_SO


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(build_oracle())
        lib.lfo_align.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(AlignOut), C.c_void_p]
        lib.lfo_ksw_extend2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 8 + [C.POINTER(C.c_int)] * 2
        lib.lfo_pack_ref.argtypes = [C.c_char_p, C.c_int64, C.c_void_p]
        lib.lfo_align_chain.argtypes = [C.POINTER(Ref), C.POINTER(Seed), C.c_int, C.c_char_p, C.c_int32, C.c_int,
                                        C.POINTER(Sam), C.c_int, C.POINTER(C.c_int), C.POINTER(Stats)]
        lib.lfo_free_sam.argtypes = [C.POINTER(Sam), C.c_int]
        _oracle = lib
    return _oracle


def have_ref():
    return os.path.exists(REF_SHIM_SO)


def ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SHIM_SO)
        lib.ref_align.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int)]
        lib.ref_extend.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 8 + [C.POINTER(C.c_int)] * 2
        lib.ref_set_index.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        lib.ref_align_chain.argtypes = [C.POINTER(Seed), C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(Sam), C.c_int, C.POINTER(C.c_int)]
        lib.ref_replay_chains.restype = C.c_double
        if hasattr(lib, "ref_fm_load"):
            lib.ref_fm_load.argtypes = [C.c_char_p, C.c_int]
            lib.ref_fm_info.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
            lib.ref_fm_seed.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int)]
            lib.ref_fm_seed_batch.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_uint32), C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
            lib.ref_fm_seed_batch.restype = C.c_double
            if hasattr(lib, "ref_fm_seed_batch_digest"):
                lib.ref_fm_seed_batch_digest.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_uint32), C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p]
                lib.ref_fm_seed_batch_digest.restype = C.c_double
        _ref = lib
    return _ref


def is_leaf(q: int, t: int) -> bool:
    """edlib's choice of full traceback over Hirschberg (lib/edlib/edlib.cpp:1117-1119)"""
    return 20 * ((q + 63) // 64) * t + 8 * t < (1 << 20) or t < 2


def oracle_align(q: bytes, t: bytes, mode: int, want_path=True):
    out = AlignOut()
    ops = C.create_string_buffer(len(q) + len(t) + 1)
    rc = oracle().lfo_align(q, len(q), t, len(t), mode, 1 if want_path else 0, C.byref(out), ops)
    assert rc == 0, rc
    return out.edit_distance, out.end_location, ops.raw[:out.n_ops]


def ref_align(q: bytes, t: bytes, mode: int, want_path=True):
    ed, end, n = C.c_int(), C.c_int(), C.c_int()
    ops = C.create_string_buffer(len(q) + len(t) + 1)
    ref().ref_align(q, len(q), t, len(t), mode, 1 if want_path else 0, C.byref(ed), C.byref(end), ops, C.byref(n))
    return ed.value, end.value, ops.raw[:n.value]


def _extend(fn, q: bytes, t: bytes, o_del, e_del, o_ins, e_ins, w, zdrop, h0=None, mat=CLIP_MAT, end_bonus=0):
    qle, tle = C.c_int(), C.c_int()
    qb, tb = C.create_string_buffer(q, len(q)), C.create_string_buffer(t, len(t))
    sc = fn(len(q), qb, len(t), tb, 5, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop,
            len(q) if h0 is None else h0, C.byref(qle), C.byref(tle))
    return sc, qle.value, tle.value


def oracle_extend(q, t, *a, **k):
    return _extend(oracle().lfo_ksw_extend2, q, t, *a, **k)


def ref_extend(q, t, *a, **k):
    return _extend(ref().ref_extend, q, t, *a, **k)


def pack_ref(seq: bytes):
    pac = C.create_string_buffer(len(seq) // 4 + 2)
    oracle().lfo_pack_ref(seq, len(seq), pac)
    return pac


class RefIndex:
    """A packed reference plus its contig table, usable by both checkers."""

    def __init__(self, seq: bytes, contig_lens=None):
        self.seq = seq
        self.pac = pack_ref(seq)
        lens = contig_lens or [len(seq)]
        assert sum(lens) == len(seq)
        self.n = len(lens)
        self.off = (C.c_int64 * self.n)()
        self.len = (C.c_int32 * self.n)(*lens)
        o = 0
        for i, l in enumerate(lens):
            self.off[i] = o
            o += l
        self.cref = Ref(C.cast(self.pac, C.c_void_p), len(seq), self.n, self.off, self.len)


class PacIndex(RefIndex):
    """RefIndex over an existing 2-bit reference (numpy uint8, bwa .pac layout) and contig table."""

    def __init__(self, pac, l_pac, contig_off, contig_len):
        self.keep = (pac, )
        self.seq = None
        self.l_pac = int(l_pac)
        self.pac = C.cast(pac.ctypes.data, C.c_void_p)
        self.n = len(contig_off)
        self.off = (C.c_int64 * self.n)(*[int(x) for x in contig_off])
        self.len = (C.c_int32 * self.n)(*[int(x) for x in contig_len])
        self.cref = Ref(self.pac, self.l_pac, self.n, self.off, self.len)


def _sam_list(arr, n):
    out = []
    for i in range(n):
        s = arr[i]
        out.append(dict(flag=s.flag, pos=s.pos, posEnd=s.posEnd, qStart=s.qStart, qEnd=s.qEnd, nm=s.nmCount,
                        cigar=C.string_at(s.cigar).decode(), md=C.string_at(s.md).decode()))
    return out


def oracle_align_chain(idx: RefIndex, seeds, query: bytes, is_rev: int):
    arr = (Seed * len(seeds))(*[Seed(*s) for s in seeds])
    sam = (Sam * 64)()
    n = C.c_int(0)
    st = Stats()
    rc = oracle().lfo_align_chain(C.byref(idx.cref), arr, len(seeds), query, len(query), is_rev, sam, 64, C.byref(n), C.byref(st))
    assert rc == 0
    out = _sam_list(sam, n.value)
    oracle().lfo_free_sam(sam, n.value)
    return out, st


def ref_align_chain(idx: RefIndex, seeds, query: bytes, is_rev: int):
    ref().ref_set_index(C.cast(idx.pac, C.c_void_p), len(idx.seq) if idx.seq is not None else idx.l_pac, idx.n, idx.off, idx.len)
    arr = (Seed * len(seeds))(*[Seed(*s) for s in seeds])
    sam = (Sam * 64)()
    n = C.c_int(0)
    ref().ref_align_chain(arr, len(seeds), query, len(query), is_rev, sam, 64, C.byref(n))
    out = _sam_list(sam, n.value)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for i in range(n.value):
        libc.free(sam[i].cigar)
        libc.free(sam[i].md)
    return out


# ---- FM-index seeding through the reference's own code (oracle/_ref/libref_shim.so) ----
import numpy as np  # noqa: E402

SEED_DT = np.dtype([("tPos", "<u4"), ("qPos", "<u4"), ("len", "<u4")])


def write_fasta(path: str, ref_ascii: np.ndarray, name: str = "chr1"):
    with open(path, "wb") as f:
        f.write(b">" + name.encode() + b"\n")
        b = np.asarray(ref_ascii, dtype=np.uint8).tobytes()
        for i in range(0, len(b), 80):
            f.write(b[i:i + 80] + b"\n")


def ref_fm_load(fasta_path: str, k_cache: int = 12):
    """bwa index + lordFAST's k-mer table, built and loaded by the reference's own code.  Returns an FmIndex-like
    namespace whose arrays are copies of the reference's in-memory arrays."""
    import types
    lib = ref()
    if lib.ref_fm_load(fasta_path.encode(), k_cache):
        raise RuntimeError("ref_fm_load failed")
    sc = (C.c_uint64 * 12)()
    bwt, sa, cache = C.c_void_p(), C.c_void_p(), C.c_void_p()
    lib.ref_fm_info(sc, C.byref(bwt), C.byref(sa), C.byref(cache))
    primary, L2, seq_len, bwt_size, n_sa, sa_intv, l_pac, kc = sc[0], list(sc[1:6]), sc[6], sc[7], sc[8], sc[9], sc[10], sc[11]
    grab = lambda p, n, dt: np.frombuffer((C.c_uint8 * (n * np.dtype(dt).itemsize)).from_address(p.value), dtype=dt).copy()
    return types.SimpleNamespace(bwt=grab(bwt, bwt_size, np.uint32), sa=grab(sa, n_sa, np.uint64), primary=int(primary),
                                 L2=np.array(L2, dtype=np.uint64), seq_len=int(seq_len), sa_intv=int(sa_intv), l_pac=int(l_pac),
                                 k_cache=int(kc), cache=grab(cache, 2 * 4 ** int(kc), np.uint64).reshape(-1, 2))


def ref_fm_seed(read: bytes, sampling_count=1000, min_anchor_len=14, max_ref_hits=1000):
    lib = ref()
    cap = sampling_count * max_ref_hits + 1
    f = np.zeros(cap, dtype=SEED_DT); r = np.zeros(cap, dtype=SEED_DT)
    nf, nr = C.c_int(), C.c_int()
    if lib.ref_fm_seed(read + b"\0", len(read), sampling_count, min_anchor_len, max_ref_hits, f.ctypes.data, C.byref(nf), r.ctypes.data, C.byref(nr)):
        raise RuntimeError("ref_fm_seed: no index loaded")
    return f[:nf.value].copy(), r[:nr.value].copy()


def seed_list_digests(seeds: np.ndarray, off: np.ndarray) -> np.ndarray:
    """per read: sum over i of (i + 1) * mix(seed i) mod 2^64 -- the digest ref_fm_seed_batch_digest computes (oracle/ref_shim.cpp)"""
    off = off.astype(np.int64)
    n = len(off) - 1
    if len(seeds) == 0:
        return np.zeros(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        mix = (seeds["tPos"].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + seeds["qPos"].astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
               + seeds["len"].astype(np.uint64) * np.uint64(0x165667B19E3779F9) + np.uint64(0x27D4EB2F165667C5))
        rank = (np.arange(len(seeds), dtype=np.int64) - np.repeat(off[:-1], np.diff(off)) + 1).astype(np.uint64)
        w = mix * rank
        cs = np.concatenate([[np.uint64(0)], np.cumsum(w, dtype=np.uint64)])
        return cs[off[1:]] - cs[off[:-1]]


def ref_fm_seed_digests(reads: np.ndarray, off: np.ndarray, sampling_count=1000, min_anchor_len=14, max_ref_hits=1000, threads=None):
    """(digests[n, 2], counts[n, 2], seconds) of the reference's seed lists (forward, reverse) for every read, on all host threads"""
    lib = ref()
    n = len(off) - 1
    rb = np.asarray(reads, dtype=np.uint8).tobytes()
    qs = [rb[int(off[i]):int(off[i + 1])] + b"\0" for i in range(n)]
    arr = (C.c_char_p * n)(*qs)
    ql = (C.c_uint32 * n)(*[len(q) - 1 for q in qs])
    dig = np.zeros((n, 2), dtype=np.uint64); cnt = np.zeros((n, 2), dtype=np.uint32)
    hits = C.c_uint64()
    sec = lib.ref_fm_seed_batch_digest(n, arr, ql, sampling_count, min_anchor_len, max_ref_hits, threads or (os.cpu_count() or 1), C.byref(hits), dig.ctypes.data, cnt.ctypes.data)
    return dig, cnt, sec
