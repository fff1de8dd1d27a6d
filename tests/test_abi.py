"""CPU tier: the C-ABI library loads without a GPU and exports every symbol the
headers declare; without a device its compute entry points fail loudly instead
of falling back to a CPU path."""
import os
import re

import pytest

from sregex_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(sre_[a-z0-9_]+)\s*\(", text))
    return sorted(n for n in names if not n.endswith(("_t", "_pt")))


def test_library_exports_every_declared_symbol(cuda):
    import ctypes
    L = ctypes.CDLL(capi.CUDA_LIB)
    names = _declared("sregex/sregex.h") + _declared("sregex_cuda.h")
    assert len(names) >= 16 + 12
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_front_end_of_product_library_matches_reference(golden, cuda):
    """parse/compile live in the product library too: same dumps."""
    n = 0
    for b in golden["blocks"][::5]:
        if "skip" in b or "error" in b:
            continue
        p = cuda.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        assert p.dump() == b["dump"]
        p.close()
        n += 1
    assert n > 300


def test_no_cpu_fallback_without_device(cuda):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    p = cuda.compile(b"a+b")
    L = cuda.L
    pool = L.sre_create_pool(1024)
    assert not L.sre_vm_thompson_create_ctx(pool, p.prog)      # NULL, not a CPU context
    assert not L.sre_vm_pike_create_ctx(pool, p.prog, None, 0)
    L.sre_destroy_pool(pool)
    from sregex_b200 import cuda as cu
    with pytest.raises(cu.SreCudaError):
        cu.CudaProgram(b"a+b")
