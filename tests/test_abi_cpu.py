"""The C-ABI library loads on a CPU-only box, exports every symbol include/dexterity_b200.h
declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dexterity_b200.h")).read()
    return sorted(set(re.findall(r"\b(dex_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from dexterity_b200 import backend
    lib = backend.lib()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dexterity_b200.h but not exported"


def test_blob_header_is_single_source_of_truth(reorient_model):
    from dexterity_b200.model import compile as C
    src = open(os.path.join(ROOT, "include", "dex_model_blob.h")).read()
    assert len(C.FIELDS) == len(re.findall(r"^\s*X\(", src, re.M))
    blob = reorient_model.to_blob()
    assert blob[:4] == (0x44455842).to_bytes(4, "little") and len(blob) % 8 == 0


def test_no_cpu_fallback(reorient_model):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from dexterity_b200 import backend
    with pytest.raises(backend.DexError, match="no CUDA device|CUDA"):
        backend.PhysicsModel(reorient_model, 0)


def test_model_validation_rejects_bad_blob():
    from dexterity_b200 import backend
    lib = backend.lib()
    out = ctypes.c_void_p()
    buf = ctypes.create_string_buffer(64)
    assert lib.dex_model_create(buf, 64, 0, ctypes.byref(out)) < 0
    assert lib.dex_last_error()
