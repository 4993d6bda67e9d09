"""The C-ABI library loads (no GPU needed) and exports every function include/difacto_b200.h declares."""
import ctypes
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def declared_functions():
    src = open(os.path.join(ROOT, "include", "difacto_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dfb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from difacto_b200 import build, capi
    build.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/difacto_b200.h but not exported: {missing}"
    # the Python binding covers the same surface
    assert sorted(capi.EXPORTS) == names, set(capi.EXPORTS) ^ set(names)


def test_host_only_entry_points_work_without_a_gpu():
    from difacto_b200 import capi
    import numpy as np
    # ps-lite range rule (postoffice.cc:127-136): pure host code in the C-ABI
    assert capi.key_owner(0, 8) == 0 and capi.key_owner(2 ** 64 - 1, 8) == 7
    w = (2 ** 64 - 1) // 8
    assert capi.key_owner(w - 1, 8) == 0 and capi.key_owner(w, 8) == 1
    keys = np.array([1, w - 1, w, 3 * w + 5, 2 ** 64 - 1], dtype=np.uint64)
    assert list(capi.shard_bounds(keys, 8)) == [0, 2, 3, 3, 4, 4, 4, 4, 5]
    # no CPU fallback: creating an engine without a device fails loudly with a CUDA error
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        try:
            capi.Engine(V_dim=4)
            raise AssertionError("engine creation must fail without a GPU")
        except capi.DfbError as e:
            assert e.code == capi.DFB_ERR_CUDA
