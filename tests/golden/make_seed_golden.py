#!/usr/bin/env python
"""Writes tests/golden/seeds_small.npz: the seed lists the REFERENCE's own getLocs_extend_whole_step (src/BWT.cpp:312-394,
through oracle/_ref/libref_shim.so: bwa index + k-mer table built by the reference's code) produces for the small case of
tests/test_seed.py, for each of its parameter sets.  Run in the build container (needs /root/reference -> oracle/_ref)."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _oracle  # noqa: E402
import test_seed  # noqa: E402

ref, reads, off = test_seed.small_case()
out = {"ref": ref, "reads": reads, "off": off}
with tempfile.TemporaryDirectory() as d:
    fa = os.path.join(d, "r.fa")
    _oracle.write_fasta(fa, ref)
    _oracle.ref_fm_load(fa, 8)
    rb = reads.tobytes()
    for k, prm in enumerate(test_seed.PARAM_SETS):
        fl, rl, fo, ro = [], [], [0], [0]
        for i in range(len(off) - 1):
            f, r = _oracle.ref_fm_seed(rb[int(off[i]):int(off[i + 1])], **prm)
            fl.append(f); rl.append(r); fo.append(fo[-1] + len(f)); ro.append(ro[-1] + len(r))
        out[f"p{k}_fwd"] = np.concatenate(fl); out[f"p{k}_rev"] = np.concatenate(rl)
        out[f"p{k}_fwd_off"] = np.array(fo, dtype=np.uint64); out[f"p{k}_rev_off"] = np.array(ro, dtype=np.uint64)
        print(prm, "forward", fo[-1], "reverse", ro[-1])
np.savez_compressed(os.path.join(HERE, "seeds_small.npz"), **out)
