"""ctypes binding of liblfgpu.so (include/lf_gpu.h) -- the host-side mirror of lordFAST's alignment
stage operators.  PyTorch is not needed here; the library owns its CUDA stream and buffers.

There is deliberately no fallback: if liblfgpu.so is missing or no CUDA device is visible the calls
raise.  `lib_path` exists so that tests can point the same binding at the test-only emulator build.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

N_CLASSES = 29  # LF_NCLS: 16 register classes, large, bad, 4 banded register classes, 7 lane-group classes
CLASS_NAMES = ([f"NW{w}{s}" for w in (1, 2, 3, 4, 6, 8, 12, 16) for s in ("", "_shw")] + ["large", "bad"]
               + [f"band{b}_NW{w}" for b, w in ((3, 6), (4, 8), (4, 12), (5, 16))]
               + ["group_path16", "group_path32", "group_path64", "group_dist32", "group_dist64", "group_dist128", "group_dist256"])
# the size-class kernels run on 18 streams; give each its own hardware queue (must be set before CUDA starts)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "liblfgpu.so")

LF_MODE_NW, LF_MODE_SHW = 0, 1
LF_F_READ_REV, LF_F_REVERSE_BOTH, LF_F_RC_QUERY, LF_F_NO_PATH = 1, 2, 4, 8
LF_MAT_CLIP, LF_MAT_DEFAULT = 0, 1

ALIGN_TASK = np.dtype([("read_id", "<u4"), ("q_off", "<u4"), ("q_len", "<u4"), ("t_off", "<u4"), ("t_len", "<u4"),
                       ("flags", "<u2"), ("mode", "u1"), ("reserved", "u1")])
ALIGN_RESULT = np.dtype([("edit_distance", "<i4"), ("end_location", "<i4"), ("ops_off", "<u8"), ("ops_len", "<u4"),
                         ("status", "<i4")])
EXTEND_TASK = np.dtype([("read_id", "<u4"), ("q_off", "<u4"), ("q_len", "<u4"), ("t_off", "<u4"), ("t_len", "<u4"),
                        ("flags", "<u2"), ("matrix", "u1"), ("reserved", "u1"), ("o_del", "<i4"), ("e_del", "<i4"),
                        ("o_ins", "<i4"), ("e_ins", "<i4"), ("w", "<i4"), ("zdrop", "<i4"), ("h0", "<i4")])
EXTEND_RESULT = np.dtype([("score", "<i4"), ("qle", "<i4"), ("tle", "<i4")])
SEED = np.dtype([("tPos", "<u4"), ("qPos", "<u4"), ("len", "<u4")])
CHAIN = np.dtype([("seed_off", "<u8"), ("n_seeds", "<u4"), ("read_id", "<u4"), ("is_rev", "<u4"), ("reserved", "<u4")])
SAM_RECORD = np.dtype([("chain_id", "<u4"), ("flag", "<u4"), ("pos", "<u4"), ("posEnd", "<u4"), ("qStart", "<u4"), ("qEnd", "<u4"),
                       ("nmCount", "<i4"), ("cigar_len", "<u4"), ("md_len", "<u4"), ("_pad", "<u4"), ("cigar_off", "<u8"), ("md_off", "<u8")])
assert SEED.itemsize == 12 and CHAIN.itemsize == 24 and SAM_RECORD.itemsize == 56
assert ALIGN_TASK.itemsize == 24 and ALIGN_RESULT.itemsize == 24 and EXTEND_TASK.itemsize == 52 and EXTEND_RESULT.itemsize == 12


class Reads(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("offsets", C.c_void_p), ("n_reads", C.c_uint32)]


class Contigs(C.Structure):
    _fields_ = [("offset", C.c_void_p), ("len", C.c_void_p), ("n", C.c_int32)]


class ChainStats(C.Structure):
    _fields_ = [("round1_tasks", C.c_uint64), ("round2_extends", C.c_uint64), ("round3_tasks", C.c_uint64), ("records", C.c_uint64),
                ("ms_tasks", C.c_float), ("ms_round1", C.c_float), ("ms_rounds23", C.c_float), ("ms_emit", C.c_float), ("ms_merge", C.c_float), ("reserved", C.c_float)]


class FmIndexC(C.Structure):   # lf_fm_index
    _fields_ = [("bwt", C.c_void_p), ("bwt_size", C.c_uint64), ("primary", C.c_uint64), ("L2", C.c_uint64 * 5), ("seq_len", C.c_uint64),
                ("sa", C.c_void_p), ("n_sa", C.c_uint64), ("sa_intv", C.c_int32), ("k_cache", C.c_int32), ("cache", C.c_void_p), ("l_pac", C.c_int64)]


class SeedParams(C.Structure):   # lf_seed_params: MIN_ANCHOR_LEN, SAMPLING_COUNT, MAX_REF_HITS (src/CommandLineParser.cpp:51-55)
    _fields_ = [("min_anchor_len", C.c_int32), ("sampling_count", C.c_int32), ("max_ref_hits", C.c_int32)]


class SeedStats(C.Structure):   # lf_seed_stats
    _fields_ = [("search_ms", C.c_float), ("locate_ms", C.c_float), ("positions", C.c_uint64), ("hits", C.c_uint64),
                ("search_steps", C.c_uint64), ("locate_steps", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("align_tasks", C.c_uint64), ("extend_tasks", C.c_uint64),
                ("cells", C.c_uint64), ("word_columns", C.c_uint64), ("last_run_ms", C.c_float),
                ("last_main_kernel_ms", C.c_float), ("last_main_word_columns", C.c_uint64)]


EXPORTS = ["lf_gpu_init", "lf_gpu_prewarm", "lf_gpu_destroy", "lf_gpu_last_error", "lf_gpu_host_alloc", "lf_gpu_host_free",
           "lf_gpu_ops_capacity", "lf_gpu_align_batch", "lf_gpu_extend_batch", "lf_gpu_upload_reads", "lf_gpu_pack_reads",
           "lf_gpu_upload_align_tasks", "lf_gpu_run_align", "lf_gpu_sync", "lf_gpu_download_align",
           "lf_gpu_upload_extend_tasks", "lf_gpu_run_extend", "lf_gpu_download_extend", "lf_gpu_get_stats",
           "lf_gpu_int32_peak", "lf_gpu_class_timeline", "lf_gpu_class_counts", "lf_gpu_align_chains", "lf_chain_results_records", "lf_chain_results_text",
           "lf_chain_results_stats", "lf_chain_results_free",
           "lf_gpu_seed_init", "lf_gpu_seed_batch", "lf_seed_results_list", "lf_seed_results_free", "lf_gpu_seed_cache_download", "lf_gpu_seed_stats"]


class LfGpuError(RuntimeError):
    pass


def load(lib_path: str | None = None) -> C.CDLL:
    path = lib_path or os.environ.get("LFGPU_LIB") or DEFAULT_LIB  # LFGPU_LIB: another CUDA build of the same library (kernel experiments)
    if not os.path.exists(path):
        raise LfGpuError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback for the alignment stage)")
    lib = C.CDLL(path)
    vp, sz = C.c_void_p, C.c_size_t
    lib.lf_gpu_init.argtypes = [C.POINTER(vp), vp, C.c_int64, vp, C.c_int]
    lib.lf_gpu_destroy.argtypes = [vp]
    lib.lf_gpu_destroy.restype = None
    lib.lf_gpu_last_error.argtypes = [vp]
    lib.lf_gpu_last_error.restype = C.c_char_p
    lib.lf_gpu_host_alloc.argtypes = [sz]
    lib.lf_gpu_host_alloc.restype = vp
    lib.lf_gpu_host_free.argtypes = [vp]
    lib.lf_gpu_host_free.restype = None
    lib.lf_gpu_ops_capacity.argtypes = [vp, sz]
    lib.lf_gpu_ops_capacity.restype = sz
    lib.lf_gpu_align_batch.argtypes = [vp, C.POINTER(Reads), vp, sz, vp, vp, sz]
    lib.lf_gpu_extend_batch.argtypes = [vp, C.POINTER(Reads), vp, sz, vp]
    lib.lf_gpu_upload_reads.argtypes = [vp, C.POINTER(Reads)]
    lib.lf_gpu_pack_reads.argtypes = [vp]
    lib.lf_gpu_upload_align_tasks.argtypes = [vp, vp, sz]
    lib.lf_gpu_run_align.argtypes = [vp]
    lib.lf_gpu_sync.argtypes = [vp]
    lib.lf_gpu_download_align.argtypes = [vp, vp, vp, sz]
    lib.lf_gpu_upload_extend_tasks.argtypes = [vp, vp, sz]
    lib.lf_gpu_run_extend.argtypes = [vp]
    lib.lf_gpu_download_extend.argtypes = [vp, vp]
    lib.lf_gpu_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.lf_gpu_int32_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    lib.lf_gpu_class_timeline.argtypes = [vp, vp, vp, C.c_int]
    lib.lf_gpu_class_counts.argtypes = [vp, vp, C.c_int]
    lib.lf_gpu_align_chains.argtypes = [vp, C.POINTER(Reads), C.POINTER(Contigs), vp, vp, sz, vp, C.POINTER(vp)]
    lib.lf_chain_results_records.argtypes = [vp, C.POINTER(sz)]
    lib.lf_chain_results_records.restype = vp
    lib.lf_chain_results_text.argtypes = [vp, C.POINTER(sz)]
    lib.lf_chain_results_text.restype = vp
    lib.lf_chain_results_stats.argtypes = [vp, C.POINTER(ChainStats)]
    lib.lf_chain_results_free.argtypes = [vp]
    lib.lf_chain_results_free.restype = None
    lib.lf_gpu_seed_init.argtypes = [vp, C.POINTER(FmIndexC)]
    lib.lf_gpu_seed_batch.argtypes = [vp, C.POINTER(Reads), C.POINTER(SeedParams), C.POINTER(vp)]
    lib.lf_seed_results_list.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(sz)]
    lib.lf_seed_results_list.restype = vp
    lib.lf_seed_results_free.argtypes = [vp]
    lib.lf_seed_results_free.restype = None
    lib.lf_gpu_seed_cache_download.argtypes = [vp, vp, sz]
    lib.lf_gpu_seed_stats.argtypes = [vp, C.POINTER(SeedStats)]
    return lib


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


class PinnedArray:
    """A numpy view over pinned host memory obtained from the library."""

    def __init__(self, lib, nbytes: int):
        self.lib, self.nbytes = lib, max(int(nbytes), 1)
        self.ptr = lib.lf_gpu_host_alloc(self.nbytes)
        if not self.ptr:
            raise LfGpuError("lf_gpu_host_alloc failed")
        self.buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)

    def view(self, dtype, count=None):
        dt = np.dtype(dtype)
        n = self.nbytes // dt.itemsize if count is None else int(count)
        return np.frombuffer(self.buf, dtype=np.uint8, count=n * dt.itemsize).view(dt)

    def free(self):
        if self.ptr:
            self.lib.lf_gpu_host_free(self.ptr)
            self.ptr = None


def decode_ops(ops: np.ndarray, off: int, n: int) -> np.ndarray:
    """Op codes (0 match, 1 insert, 2 delete, 3 mismatch) of one task out of the 2-bit stream."""
    p = np.arange(off, off + n, dtype=np.int64)
    return ((ops[p >> 2] >> ((p & 3) << 1).astype(np.uint8)) & 3).astype(np.uint8)


class LfGpu:
    """One context: the 2-bit reference resident on the device(s) + the batched alignment calls."""

    def __init__(self, pac: np.ndarray, l_pac: int, devices=None, lib_path: str | None = None):
        self.lib = load(lib_path)
        self.pac = np.ascontiguousarray(pac, dtype=np.uint8)
        if len(self.pac) < l_pac // 4 + 1:
            raise LfGpuError("pac shorter than l_pac/4+1 bytes")
        self.ctx = C.c_void_p()
        dev = None
        ndev = 0
        if devices:
            dev = np.asarray(devices, dtype=np.int32)
            ndev = len(dev)
        rc = self.lib.lf_gpu_init(C.byref(self.ctx), _ptr(self.pac), int(l_pac), _ptr(dev) if ndev else None, ndev)
        if rc != 0:
            self.ctx = C.c_void_p()
            raise LfGpuError(f"lf_gpu_init failed with status {rc} (no CUDA device? there is no CPU fallback)")
        self._keep = []

    def close(self):
        if self.ctx:
            self.lib.lf_gpu_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise LfGpuError(f"{what} failed with status {rc}: {self.lib.lf_gpu_last_error(self.ctx).decode()}")

    def _reads(self, bases, offsets: np.ndarray) -> Reads:
        """bases = None: the reads the last upload_reads / seed_batch left on the device (lf_gpu_align_chains accepts that)"""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        if bases is None:
            self._keep = [offsets]
            return Reads(None, _ptr(offsets), len(offsets) - 1)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        self._keep = [bases, offsets]
        return Reads(_ptr(bases), _ptr(offsets), len(offsets) - 1)

    # ---- one-call forms (host buffers in, host buffers out) ----
    def align_batch(self, bases, offsets, tasks: np.ndarray):
        tasks = np.ascontiguousarray(tasks, dtype=ALIGN_TASK)
        n = len(tasks)
        res = np.zeros(n, dtype=ALIGN_RESULT)
        cap = self.lib.lf_gpu_ops_capacity(_ptr(tasks), n)
        ops = np.zeros(cap, dtype=np.uint8)
        r = self._reads(bases, offsets)
        self._check(self.lib.lf_gpu_align_batch(self.ctx, C.byref(r), _ptr(tasks), n, _ptr(res), _ptr(ops), cap), "lf_gpu_align_batch")
        return res, ops

    def extend_batch(self, bases, offsets, tasks: np.ndarray):
        tasks = np.ascontiguousarray(tasks, dtype=EXTEND_TASK)
        n = len(tasks)
        res = np.zeros(n, dtype=EXTEND_RESULT)
        r = self._reads(bases, offsets)
        self._check(self.lib.lf_gpu_extend_batch(self.ctx, C.byref(r), _ptr(tasks), n, _ptr(res)), "lf_gpu_extend_batch")
        return res

    def align_chains(self, bases, offsets, contig_off, contig_len, seeds: np.ndarray, chains: np.ndarray, want_text=True):
        """alignChain_edlib for a chunk of chains.  Returns (records, text bytes, ChainStats)."""
        seeds = np.ascontiguousarray(seeds, dtype=SEED)
        chains = np.ascontiguousarray(chains, dtype=CHAIN)
        co = np.ascontiguousarray(contig_off, dtype=np.int64)
        cl = np.ascontiguousarray(contig_len, dtype=np.int32)
        r = self._reads(bases, offsets)
        cg = Contigs(_ptr(co), _ptr(cl), len(co))
        out = C.c_void_p()
        self._check(self.lib.lf_gpu_align_chains(self.ctx, C.byref(r), C.byref(cg), _ptr(seeds), _ptr(chains), len(chains),
                                                 _ptr(self.pac), C.byref(out)), "lf_gpu_align_chains")
        n, nb = C.c_size_t(), C.c_size_t()
        rp = self.lib.lf_chain_results_records(out, C.byref(n))
        tp = self.lib.lf_chain_results_text(out, C.byref(nb))
        recs = np.frombuffer((C.c_uint8 * (n.value * SAM_RECORD.itemsize)).from_address(rp), dtype=SAM_RECORD).copy() if n.value else np.zeros(0, dtype=SAM_RECORD)
        text = bytes((C.c_uint8 * nb.value).from_address(tp)) if (want_text and nb.value) else b""
        st = ChainStats()
        self.lib.lf_chain_results_stats(out, C.byref(st))
        self.lib.lf_chain_results_free(out)
        return recs, text, st

    # ---- FM-index seeding (getLocs_extend_whole_step for a batch of reads) ----
    def seed_init(self, fm):
        """fm: lordfast_b200.fmindex.FmIndex (or anything with its fields) holding bwa's arrays."""
        bwt = np.ascontiguousarray(fm.bwt, dtype=np.uint32)
        sa = np.ascontiguousarray(fm.sa, dtype=np.uint64)
        cache = None if fm.cache is None else np.ascontiguousarray(fm.cache, dtype=np.uint64)
        c = FmIndexC(_ptr(bwt), len(bwt), int(fm.primary), (C.c_uint64 * 5)(*[int(x) for x in fm.L2]), int(fm.seq_len), _ptr(sa), len(sa),
                     int(fm.sa_intv), int(fm.k_cache), None if cache is None else _ptr(cache), int(fm.l_pac))
        self._check(self.lib.lf_gpu_seed_init(self.ctx, C.byref(c)), "lf_gpu_seed_init")
        self._k_cache = int(fm.k_cache)

    def seed_batch(self, bases, offsets, min_anchor_len=14, sampling_count=1000, max_ref_hits=1000, copy=True):
        """Returns (fwd seeds, fwd offsets, rev seeds, rev offsets); offsets have n_reads + 1 entries.  bases = None: the resident reads."""
        prm = SeedParams(min_anchor_len, sampling_count, max_ref_hits)
        out = C.c_void_p()
        if bases is None:
            rc = self.lib.lf_gpu_seed_batch(self.ctx, None, C.byref(prm), C.byref(out))
        else:
            r = self._reads(bases, offsets)
            rc = self.lib.lf_gpu_seed_batch(self.ctx, C.byref(r), C.byref(prm), C.byref(out))
        self._check(rc, "lf_gpu_seed_batch")
        res = []
        nr = None
        for rev in (0, 1):
            op, n = C.c_void_p(), C.c_size_t()
            lp = self.lib.lf_seed_results_list(out, rev, C.byref(op), C.byref(n))
            if nr is None:
                nr = len(offsets) - 1 if offsets is not None else self._last_n_reads
            off = np.frombuffer((C.c_uint8 * ((nr + 1) * 8)).from_address(op.value), dtype=np.uint64)
            lst = np.frombuffer((C.c_uint8 * (n.value * SEED.itemsize)).from_address(lp), dtype=SEED) if n.value else np.zeros(0, dtype=SEED)
            res += [lst.copy() if copy else lst, off.copy() if copy else off]
        self._last_n_reads = nr
        self.lib.lf_seed_results_free(out)
        return tuple(res)

    def seed_cache(self) -> np.ndarray:
        n = 4 ** self._k_cache
        out = np.zeros((n, 2), dtype=np.uint64)
        self._check(self.lib.lf_gpu_seed_cache_download(self.ctx, _ptr(out), n), "lf_gpu_seed_cache_download")
        return out

    def seed_stats(self) -> dict:
        st = SeedStats()
        self.lib.lf_gpu_seed_stats(self.ctx, C.byref(st))
        return {k: getattr(st, k) for k, _ in SeedStats._fields_}

    # ---- phased forms (keep a batch resident in HBM) ----
    def upload_reads(self, bases, offsets):
        r = self._reads(bases, offsets)
        self._check(self.lib.lf_gpu_upload_reads(self.ctx, C.byref(r)), "lf_gpu_upload_reads")

    def upload_align_tasks(self, tasks: np.ndarray):
        assert tasks.dtype == ALIGN_TASK and tasks.flags["C_CONTIGUOUS"]
        self._check(self.lib.lf_gpu_upload_align_tasks(self.ctx, _ptr(tasks), len(tasks)), "lf_gpu_upload_align_tasks")

    def pack_reads(self):
        self._check(self.lib.lf_gpu_pack_reads(self.ctx), "lf_gpu_pack_reads")

    def run_align(self):
        self._check(self.lib.lf_gpu_run_align(self.ctx), "lf_gpu_run_align")

    def sync(self):
        self._check(self.lib.lf_gpu_sync(self.ctx), "lf_gpu_sync")

    def download_align(self, res: np.ndarray, ops: np.ndarray):
        self._check(self.lib.lf_gpu_download_align(self.ctx, _ptr(res), _ptr(ops), ops.nbytes), "lf_gpu_download_align")

    def stats(self) -> Stats:
        s = Stats()
        self._check(self.lib.lf_gpu_get_stats(self.ctx, C.byref(s)), "lf_gpu_get_stats")
        return s

    def class_timeline(self):
        a = np.zeros(N_CLASSES, dtype=np.float32); b = np.zeros(N_CLASSES, dtype=np.float32)
        self._check(self.lib.lf_gpu_class_timeline(self.ctx, _ptr(a), _ptr(b), N_CLASSES), "lf_gpu_class_timeline")
        return a, b

    def class_counts(self) -> dict:
        """tasks per size class of the last run_align (names: api.CLASS_NAMES)"""
        a = np.zeros(N_CLASSES, dtype=np.uint32)
        self._check(self.lib.lf_gpu_class_counts(self.ctx, _ptr(a), N_CLASSES), "lf_gpu_class_counts")
        return {CLASS_NAMES[c]: int(a[c]) for c in range(N_CLASSES) if a[c]}

    def int32_peak(self, which: int) -> float:
        v = C.c_double()
        self._check(self.lib.lf_gpu_int32_peak(self.ctx, which, C.byref(v)), "lf_gpu_int32_peak")
        return v.value


def records_to_dicts(recs: np.ndarray, text: bytes):
    """lf_sam_record rows -> the dict form the tests compare with the reference's Sam_t dumps."""
    out = []
    for r in recs:
        out.append(dict(chain=int(r["chain_id"]), flag=int(r["flag"]), pos=int(r["pos"]), posEnd=int(r["posEnd"]), qStart=int(r["qStart"]),
                        qEnd=int(r["qEnd"]), nm=int(r["nmCount"]),
                        cigar=text[int(r["cigar_off"]):int(r["cigar_off"]) + int(r["cigar_len"])].decode(),
                        md=text[int(r["md_off"]):int(r["md_off"]) + int(r["md_len"])].decode()))
    return out


def workload_chains(w):
    """sim.Workload -> (seeds, chains) arrays for LfGpu.align_chains (one chain per read)."""
    seeds = np.zeros(len(w.seeds), dtype=SEED)
    seeds["tPos"], seeds["qPos"], seeds["len"] = w.seeds[:, 0], w.seeds[:, 1], w.seeds[:, 2]
    chains = np.zeros(w.n_reads, dtype=CHAIN)
    chains["seed_off"] = w.seed_off[:-1]
    chains["n_seeds"] = np.diff(w.seed_off)
    chains["read_id"] = np.arange(w.n_reads)
    chains["is_rev"] = w.is_rev
    return seeds, chains
