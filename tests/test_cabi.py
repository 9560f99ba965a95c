"""CPU: the C-ABI library loads, exports every symbol include/mlcg.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "mlcg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mlcg_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as g
    g.build()
    from ml_conformer_generator_b200 import _lib
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(_lib.EXPORTS) == names
    assert b"sm_100a" in lib.mlcg_version()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from ml_conformer_generator_b200 import _lib
    from ml_conformer_generator_b200.engine import Engine, MlcgError
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.mlcg_create(C.byref(h), 0, 2) == -3  # MLCG_E_NO_DEVICE
    with pytest.raises(MlcgError):
        Engine(torch.device("cuda:0"), "bf16")
    with pytest.raises(MlcgError):
        Engine(torch.device("cpu"), "bf16")
