"""The C-ABI library loads on a CPU-only box and exports every symbol include/vla_b200.h declares; without a
CUDA device the product path fails loudly (there is no CPU fallback).  No compute is called here."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vla_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vla_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from vla_adapter_b200 import _lib

    names = _declared()
    assert len(names) >= 18
    assert not _lib.MISSING
    for n in names:
        assert hasattr(lib, n), f"libvla_b200.so does not export {n}"
        assert n in _lib.SIGNATURES, f"ctypes binding lacks {n}"
    assert sorted(_lib.SIGNATURES) == names, "binding lists symbols the header does not declare"


def test_cfg_struct_matches_header():
    from vla_adapter_b200 import _lib

    src = open(os.path.join(ROOT, "include", "vla_b200.h")).read()
    body = src[src.index("typedef struct vla_cfg {"):src.index("} vla_cfg;")]
    fields = re.findall(r"int32_t\s+(\w+);", body)
    assert fields == [f[0] for f in _lib.VlaCfg._fields_]
    assert C.sizeof(_lib.VlaCfg) == 4 * len(fields)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from vla_adapter_b200 import _lib
    from vla_adapter_b200.engine import VLAEngine

    cfg = _lib.VlaCfg(2, 8, 7, 8, 0, 24, 27, 24, 151936, 1, 64, 1)
    h = C.c_void_p()
    assert lib.vla_create(C.byref(cfg), C.byref(h)) == -4  # VLA_ERR_CUDA
    assert not h.value
    with pytest.raises(RuntimeError):
        VLAEngine()
    from vla_adapter_b200 import ops

    with pytest.raises(RuntimeError):
        ops.linear(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from vla_adapter_b200 import _lib

    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setenv("VLA_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
