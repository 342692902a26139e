"""The C ABI: the library loads without a GPU, exports every symbol include/ance_b200.h declares, and the
host-only entry points behave.  No GPU compute is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ance_b200 import _lib
from ance_b200.search import merge_topk_host
from oracle import flat_ip_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported_and_bound(lib):
    hdr = open(os.path.join(ROOT, "include", "ance_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ance_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ance_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes SIGNATURES and the header disagree"


def test_version_and_error_string(lib):
    assert b"sm_100a" in lib.ance_version()
    assert lib.ance_launch_count() >= 0


def test_argument_errors_do_not_need_a_gpu(lib):
    assert lib.ance_index_search(None, None, 1, 1, None, None, 0, None) == 1  # ANCE_ERR_INVALID
    assert b"null handle" in lib.ance_last_error()
    assert lib.ance_index_create(7, 10, 1, C.byref(C.c_void_p())) == 1
    assert b"multiple of 8" in lib.ance_last_error()


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.ance_index_create(64, 10, 1, C.byref(h)) == 2  # ANCE_ERR_CUDA
    from ance_b200.search import IndexFlatIP
    with pytest.raises(_lib.AnceError):
        IndexFlatIP(64)


@pytest.mark.parametrize("W,k", [(1, 5), (4, 10), (8, 200), (3, 1)])
def test_merge_topk_host_matches_oracle(W, k):
    rng = np.random.default_rng(W * 100 + k)
    n, nq, d = 777, 33, 16
    P = rng.standard_normal((n, d)).astype(np.float32)
    P[300:305] = P[0:5]  # ties across shards
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    Q[0] = P[2]
    order = np.concatenate([np.arange(r, n, W) for r in range(W)])
    Pm = P[order]
    Dg, Ig = flat_ip_oracle.search_bruteforce(Pm, Q, k)
    Ds, Is, off = [], [], 0
    for r in range(W):
        m = len(range(r, n, W))
        d_, i_ = flat_ip_oracle.search_bruteforce(Pm[off:off + m], Q, k)
        Ds.append(d_)
        Is.append(np.where(i_ >= 0, i_ + off, -1))
        off += m
    Dm, Im = merge_topk_host(Ds, Is, k, n_threads=3)
    assert (Im == Ig).all() and (Dm == Dg).all()
    Do, Io = flat_ip_oracle.merge_shards(Ds, Is, k)
    assert (Im == Io).all() and (Dm == Do).all()


def test_merge_topk_host_padding_and_errors():
    D = [np.array([[3.0, 1.0, np.finfo(np.float32).min]], dtype=np.float32)]
    I = [np.array([[5, 9, -1]], dtype=np.int64)]
    Dm, Im = merge_topk_host(D + D, [I[0], I[0] + np.array([[10, 10, 0]])], 3)
    assert Im.tolist() == [[5, 15, 9]] and Dm.tolist() == [[3.0, 3.0, 1.0]]
    Dm, Im = merge_topk_host(D, I, 3)
    assert Im.tolist() == [[5, 9, -1]] and Dm[0, 2] == np.finfo(np.float32).min
    with pytest.raises(ValueError):
        merge_topk_host(D, I, 4)
