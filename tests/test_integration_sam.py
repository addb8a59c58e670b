"""Whole-program parity: the reference binary (oracle/_ref/lordfast, unmodified, CPU) against the same program with
its alignment stage replaced by liblfgpu.so through integration/lordfast_gpu_glue.cpp.  The SAM files must hold the
same records byte for byte (POS, CIGAR, NM, AS, MAPQ, MD, SA ...); the order of reads in the file depends on thread
scheduling in the reference too, so the files are compared sorted, minus the @PG line (it holds the command line).

The binaries are built in this container (where /root/reference exists) and travel to the GPU box; nothing here reads
/root/reference at run time."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "lordfast")
GPU_BIN = os.path.join(ROOT, "integration", "_build", "lordfast_gpu")
EMU_BIN = os.path.join(ROOT, "integration", "_build", "lordfast_gpu_emu")


def _dataset(tmp, **kw):
    args = [sys.executable, os.path.join(ROOT, "integration", "make_dataset.py"), str(tmp)]
    for k, v in kw.items():
        args += ["--" + k.replace("_", "-")] + [str(x) for x in (v if isinstance(v, (list, tuple)) else [v])]
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    subprocess.check_call([REF_BIN, "--index", "ref.fa"], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _sam(binary, tmp, out, threads, extra=()):
    subprocess.check_call([binary, "--search", "ref.fa", "--seq", "reads.fa", "-t", str(threads), "-o", out, *extra], cwd=tmp,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    lines = [l for l in open(os.path.join(tmp, out)) if not l.startswith("@PG")]
    return sorted(lines)


def _compare(binary, tmp, threads, extra=()):
    a = _sam(REF_BIN, tmp, "cpu.sam", threads, extra)
    b = _sam(binary, tmp, "gpu.sam", threads, extra)
    assert len(a) == len(b)
    bad = [i for i, (x, y) in enumerate(zip(a, b)) if x != y]
    assert not bad, "first differing record:\n%s\n%s" % (a[bad[0]][:400], b[bad[0]][:400])
    recs = [l.split("\t") for l in a if not l.startswith("@")]
    return recs


@pytest.mark.skipif(not os.path.isdir("/root/reference") and not os.path.exists(EMU_BIN), reason="glue binaries are built where the reference sources are")
def test_glue_sam_identical_on_emulator(tmp_path):
    """Glue logic (phasing, fine mode, scoring, sort, SAM writer) on CPU: the test-only emulator build of the library."""
    from _common import build_emu
    build_emu()
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "integration"), "emu"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    _dataset(tmp_path, ref_len=300_000, reads=30, read_len=4000, sv_frac=0.4, dups=6, contigs=2, seed=1)
    recs = _compare(EMU_BIN, tmp_path, 2)
    flags = [int(r[1]) for r in recs]
    assert any(f & 256 for f in flags), "fine mode (secondary records) not exercised"
    assert any(f & 2048 for f in flags), "split chains (supplementary records) not exercised"
    assert any(f & 16 for f in flags) and any(not (f & 16) for f in flags)


@pytest.mark.skipif(not os.path.isdir("/root/reference") and not os.path.exists(EMU_BIN), reason="glue binaries are built where the reference sources are")
def test_glue_unmapped_and_short_reads_on_emulator(tmp_path):
    """Reads below --minReadLen, reads without a candidate window and a homopolymer go through the unmapped branches of the
    glue (mapSeq :484-496, :515-524) and must print exactly as the reference prints them."""
    import numpy as np
    from _common import build_emu
    build_emu()
    _dataset(tmp_path, ref_len=300_000, reads=20, read_len=3000, sv_frac=0.3, seed=9)
    rng = np.random.default_rng(5)
    with open(os.path.join(tmp_path, "reads.fa"), "a") as f:
        for name, n in (("short1", 300), ("short2", 999), ("junk1", 5000), ("junk2", 2500)):
            f.write(">%s\n%s\n" % (name, "".join(rng.choice(list("ACGT"), size=n))))
        f.write(">polyA\n" + "A" * 3000 + "\n")
    for threads in (1, 3):
        recs = _compare(EMU_BIN, tmp_path, threads)
        assert sum(1 for r in recs if int(r[1]) == 4) >= 4


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw,threads,extra", [
    ("config1", dict(ref_len=1_000_000, reads=200, read_len=10_000, err=[0.15, 0.15], seed=1), 4, ()),
    ("dups_fine_mode", dict(ref_len=2_000_000, reads=300, read_len=8_000, err=[0.12, 0.15], seed=2, sv_frac=0.3, dups=20, contigs=3), 8, ()),
    ("clasp_chaining", dict(ref_len=1_000_000, reads=100, read_len=6_000, err=[0.12, 0.15], seed=4, sv_frac=0.3, dups=8), 4, ("--chainAlg", "clasp")),
])
def test_lordfast_gpu_sam_identical(tmp_path, name, kw, threads, extra):
    """BASELINE configs[0] and two harder variants through the real library on the B200."""
    if not (os.path.exists(GPU_BIN) and os.path.exists(REF_BIN)):
        pytest.fail("integration/_build/lordfast_gpu or oracle/_ref/lordfast missing: run __graft_entry__.build() where /root/reference exists")
    _dataset(tmp_path, **kw)
    recs = _compare(GPU_BIN, tmp_path, threads, extra)
    assert len(recs) >= kw["reads"]


@pytest.mark.gpu
def test_lordfast_gpu_sam_identical_config2_full(tmp_path):
    """BASELINE configs[1] at full size (4.6 Mbp reference, 20 000 x 10 kbp reads, 10 % SV mix) through the whole program:
    every SAM record of the GPU build against the reference binary's (20 000+ records)."""
    if not (os.path.exists(GPU_BIN) and os.path.exists(REF_BIN)):
        pytest.fail("integration/_build/lordfast_gpu or oracle/_ref/lordfast missing: run __graft_entry__.build() where /root/reference exists")
    _dataset(tmp_path, ref_len=4_600_000, reads=20_000, read_len=10_000, err=[0.12, 0.15], seed=100, sv_frac=0.10)
    recs = _compare(GPU_BIN, tmp_path, os.cpu_count() or 4)
    assert len(recs) >= 20_000


def _n_gpus():
    try:
        return len(subprocess.check_output(["nvidia-smi", "-L"], text=True).strip().splitlines())
    except Exception:
        return 0


@pytest.mark.gpu
def test_lordfast_gpu_multi_device_sam_identical(tmp_path, monkeypatch):
    """LF_GPU_DEVICES=0,1,...: one context over every GPU of the box; lf_gpu_align_chains shards the chunk's chains by
    contiguous ranges over a lane per device (every device holds the reference) and merges the records in chain order
    (the data-parallel loop of src/LordFAST.cpp:295-316).  Needs two GPUs (`gpurun --gpus 2`); skipped on one."""
    n = _n_gpus()
    if n < 2:
        pytest.skip("one GPU visible")
    if not (os.path.exists(GPU_BIN) and os.path.exists(REF_BIN)):
        pytest.fail("integration/_build/lordfast_gpu or oracle/_ref/lordfast missing")
    monkeypatch.setenv("LF_GPU_DEVICES", ",".join(str(i) for i in range(n)))
    _dataset(tmp_path, ref_len=2_000_000, reads=1500, read_len=8_000, err=[0.12, 0.15], seed=2, sv_frac=0.2, dups=20, contigs=3)
    recs = _compare(GPU_BIN, tmp_path, 8)
    assert len(recs) >= 1500
