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
    assert L.xgm_abi_version() == 2


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


def test_value_keys_are_order_preserving_and_invertible():
    """xgm_value_key / xgm_value_key_bytes / xgm_sort_key_bytes are host-only helpers (no device)."""
    import random
    rng = random.Random(11)
    vals = [bytes(rng.randrange(1, 256) for _ in range(rng.randrange(1, 9))) for _ in range(500)]
    vals += [b"", b"\x80", b"\xff", b"\xc0\x46\x40", b"\xc0\x46\x40\x01"]
    keys = [xgm.value_key(v) for v in vals]
    assert all(ex for _, ex in keys)
    for v, (k, _) in zip(vals, keys):
        assert xgm.value_key_bytes(k) == v
    order_bytes = sorted(range(len(vals)), key=lambda i: vals[i])
    order_keys = sorted(range(len(vals)), key=lambda i: (keys[i][0], vals[i]))
    assert [vals[i] for i in order_bytes] == [vals[i] for i in order_keys]
    assert xgm.value_key(b"123456789")[1] is False and xgm.value_key(b"ab\x00")[1] is False
    # Multi_MultiValueKeyMaker::operator() for one reverse SerialiseKey slot (keymaker.cc:733-743)
    assert xgm.sort_key_bytes(xgm.value_key(b"\xc3\x3a\x0e\x50")[0], True) == bytes.fromhex("3cc5f1afffff")
    assert xgm.sort_key_bytes(xgm.value_key(b"\xc0\x46\x40")[0], False) == b"\xc0\x46\x40"
