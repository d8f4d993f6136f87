"""The C-ABI library loads without a GPU and exports every symbol include/ccdm_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ccdm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ccdm_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_header_symbols_exported_and_bound(precision):
    """Both builds of the library (csrc/ptx.cuh: bf16 storage and the binary16 / TF32-class tier) export the whole header."""
    from ccdm_b200 import _lib
    handle = _lib.lib(precision)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert handle.ccdm_version() >= 100


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_struct_mirrors_match_library(precision):
    from ccdm_b200 import _lib
    handle = _lib.lib(precision)
    for which, struct in enumerate((_lib.TapGemmArgs, _lib.View, _lib.StepArgs, _lib.QSampleArgs, _lib.LossArgs)):
        assert handle.ccdm_struct_size(which) == ctypes.sizeof(struct)


def test_bad_arguments_are_rejected_without_a_gpu():
    from ccdm_b200 import _lib
    handle = _lib.lib()
    a = _lib.TapGemmArgs()
    assert handle.ccdm_tapgemm(ctypes.byref(a), None) == -1          # CCDM_ERR_BAD_ARG
    assert b"n_src" in handle.ccdm_last_error()
    assert handle.ccdm_attention_small(1, 1, 1, 16, 4, 24, 1.0, None) == -2     # CCDM_ERR_UNSUPPORTED_SHAPE


def test_no_cpu_fallback():
    import torch
    import ccdm_b200
    net = ccdm_b200.Unet(dim=32, dim_mults=(1, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 3, 8, 8), torch.zeros(1, dtype=torch.long), torch.zeros(1, 128), cond_drop_prob=0.0)
    # the product package never imports the oracle
    import sys
    import subprocess
    code = "import sys, ccdm_b200, ccdm_b200.diffusion, ccdm_b200.engine; assert not any(m.startswith('oracle') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], cwd=ROOT, check=True)
