"""The C-ABI library loads (no compute without a GPU) and exports every function include/lf_gpu.h declares."""
import ctypes
import os
import re
import subprocess

from lordfast_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    text = open(os.path.join(ROOT, "include", "lf_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lf_(?:gpu|chain_results|seed_results)_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared() == sorted(api.EXPORTS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(api.DEFAULT_LIB):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "lordfast_b200", "csrc")])
    lib = ctypes.CDLL(api.DEFAULT_LIB)
    for name in declared():
        assert hasattr(lib, name), name


def test_struct_sizes_match_header():
    assert api.ALIGN_TASK.itemsize == 24 and api.ALIGN_RESULT.itemsize == 24
    assert api.EXTEND_TASK.itemsize == 52 and api.EXTEND_RESULT.itemsize == 12


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product library must refuse to initialise (here: no GPU)."""
    import numpy as np
    import pytest
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    if not os.path.exists(api.DEFAULT_LIB):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "lordfast_b200", "csrc")])
    with pytest.raises(api.LfGpuError):
        api.LfGpu(np.zeros(64, dtype=np.uint8), 100)
