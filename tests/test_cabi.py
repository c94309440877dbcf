"""CPU: the C-ABI library builds, loads and exports every symbol include/mggan_b200.h declares,
and the ctypes prototypes in mggan/cuda_ext.py agree with the header (no compute calls here)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "mg-gan_b200"))

import header_signatures  # noqa: E402


def _lib():
    import build_ext
    return ctypes.CDLL(build_ext.build())


def test_library_exports_every_declared_symbol():
    lib = _lib()
    decl = header_signatures.parse()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/mggan_b200.h but not exported"
    lib.mggan_version.restype = ctypes.c_int
    assert lib.mggan_version() >= 100


def test_ctypes_prototypes_match_header():
    from mggan import cuda_ext
    decl = header_signatures.parse()
    for name, (ret, codes) in decl.items():
        if name in cuda_ext.SIGNATURES:
            assert cuda_ext.SIGNATURES[name] == codes, name
        else:
            assert name in cuda_ext._PLAIN and cuda_ext._PLAIN[name][0] == codes, name
    assert set(cuda_ext.EXPORTED) == set(decl)


def test_product_fails_loudly_without_device():
    """No CPU fallback: calling a kernel without CUDA raises instead of computing elsewhere."""
    import pytest
    import torch
    from mggan import cuda_ext, kernels
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cuda_ext.MgganCudaError):
        kernels.linear(torch.zeros(2, 4), torch.zeros(3, 4))


def test_product_does_not_import_oracle():
    import re
    pkg = os.path.join(ROOT, "mg-gan_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+(mggan_oracle|refshim|oracle)\b", src, re.M), (dp, f)
