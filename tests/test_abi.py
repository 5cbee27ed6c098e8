"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports every
symbol include/hyperdb_b200.h declares, the ctypes table covers them all, and compute entry points
fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hyperdb_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hdb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from hyperdb_b200 import _native as N
    names = declared_functions()
    assert len(names) >= 20
    assert sorted(N.SIGNATURES) == names
    import shutil
    if shutil.which("nvcc") or not os.path.exists(N.library_path()):
        import __graft_entry__ as ge
        ge.build()                     # incremental make: a stale library never reaches the GPU box
    handle = ctypes.CDLL(N.library_path())
    for name in names:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    N.lib()
    assert N.lib().hdb_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import hyperdb_b200
    from hyperdb_b200 import _native as N
    with pytest.raises(N.NativeError):
        hyperdb_b200.cosine_similarity(np.eye(3), np.ones(3))
    with pytest.raises(N.NativeError):
        hyperdb_b200.hyperDB_ranking_algorithm_sort(np.eye(3), np.ones(3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "local-hyperdb_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
