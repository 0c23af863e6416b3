"""CPU tests: libxgm.so loads and exports every symbol include/xgm.h declares; no CPU fallback."""
import ctypes
import os
import re

from xapiand_b200 import xgm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "xgm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xgm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = xgm.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"libxgm.so does not export {n}"
    assert set(names) == set(xgm.EXPORTS), set(names) ^ set(xgm.EXPORTS)
    assert L.xgm_abi_version() == 1


def test_no_cpu_fallback_without_device():
    """Without a CUDA device index construction must fail loudly (XGM_E_NODEVICE), never fall back."""
    import torch
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    st = xgm.lib().xgm_index_build_synthetic(100, 10, 1, 1, 0, 0, 0, 1, ctypes.byref(h))
    assert st in (xgm.E_NODEVICE, xgm.E_CUDA)
    assert b"no CPU path" in xgm.lib().xgm_last_error() or st == xgm.E_CUDA


def test_round_estimate_python_matches_oracle():
    from oracle import oracle as O
    import random
    rng = random.Random(5)
    for _ in range(2000):
        m = rng.randrange(0, 100000)
        M = m + rng.randrange(0, 200000)
        e = rng.randrange(m, M + 1)
        assert xgm.round_estimate(m, M, e) == O.round_estimate(m, M, e)
