"""GPU suite (-m gpu): the bit-exact parallel float-chain evaluator (gbrl_b200/csrc/chain.cuh) against a plain sequential
fp32 accumulation (numpy cumsum in float32 IS the sequential chain) on adversarial inputs: zero-drift random walks
(the running sum keeps crossing binades), drifting sums, wide dynamic range, many exact rounding ties, sparse members,
ragged thread partitions that start mid-row, non-finite values."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _seq_partials(mat, D, T, mode, mean):
    """math_ops.cpp:255-300 / 461-513 restated: thread t sums elements [t*ept, (t+1)*ept) (last thread to the end) of the
    flattened row-major matrix, one sequential chain per column."""
    flat = mat.reshape(-1).astype(np.float32)
    ne = flat.size
    ept = ne // T
    out = np.zeros((T, D), np.float32)
    cen = flat.copy()
    for t in range(T):
        s, e = t * ept, (ne if t == T - 1 else (t + 1) * ept)
        for col in range(D):
            first = s + ((col - s) % D)
            x = flat[first:e:D]
            if mode == 1:
                c = (x - np.float32(mean[col])).astype(np.float32)
                cen[first:e:D] = c
                x = (c * c).astype(np.float32)
            if x.size:
                with np.errstate(all="ignore"):
                    out[t, col] = np.cumsum(x, dtype=np.float32)[-1]
    return out, cen


def _gpu_partials(mat, D, T, mode, mean, impl=0, info=None):
    from gbrl_b200 import _capi
    L = _capi.lib()
    flat = np.ascontiguousarray(mat.reshape(-1), np.float32)
    part = np.zeros(T * D, np.float32)
    cen = np.zeros_like(flat)
    mean = np.ascontiguousarray(mean if mean is not None else np.zeros(D), np.float32)
    fp = C.POINTER(C.c_float)
    _capi.check(L.gbrl_b200_diag_chain_sums(flat.ctypes.data_as(fp), flat.size, D, T, mode, mean.ctypes.data_as(fp),
                                            part.ctypes.data_as(fp), cen.ctypes.data_as(fp), impl,
                                            info.ctypes.data_as(C.POINTER(C.c_double)) if info is not None else None))
    return part.reshape(T, D), cen


def _data(kind, n, D, rng):
    if kind == "walk":
        x = rng.standard_normal((n, D))
    elif kind == "drift":
        x = rng.standard_normal((n, D)) + 0.3
    elif kind == "range":
        x = rng.standard_normal((n, D)) * np.exp(5.0 * rng.standard_normal((n, D)))
    elif kind == "ties":
        x = np.round(rng.standard_normal((n, D)) * 64) / 64 + 0.125
    elif kind == "sparse":
        x = rng.standard_normal((n, D)) - 0.2
        x[rng.random((n, D)) < 0.7] = 0.0
    elif kind == "tiny":
        x = rng.standard_normal((n, D)) * 1e-36
    else:
        raise ValueError(kind)
    return x.astype(np.float32)


@pytest.mark.parametrize("kind", ["walk", "drift", "range", "ties", "sparse", "tiny"])
@pytest.mark.parametrize("D,T,n", [(1, 1, 70001), (1, 16, 200003), (2, 3, 50001), (3, 5, 33333), (4, 7, 41111), (6, 4, 9000)])
def test_parallel_chain_is_bit_identical_to_sequential(kind, D, T, n):
    rng = np.random.default_rng(hash((kind, D, T)) % (2 ** 32))
    mat = _data(kind, n, D, rng)
    for mode in (0, 1):
        mean = mat.mean(axis=0).astype(np.float32) if mode == 1 else None
        want, want_c = _seq_partials(mat, D, T, mode, mean)
        # impl 0: summaries by the whole GPU + one walking warp per chain (the product path); 2: one CTA per chain
        # (D <= 4); 3: one warp per chain
        for impl in (0, 2, 3):
            if impl == 2 and D > 4:
                continue
            got, got_c = _gpu_partials(mat, D, T, mode, mean, impl=impl)
            assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), (kind, D, T, mode, impl, want, got)
            if mode == 1:
                assert np.array_equal(want_c.view(np.uint32), got_c.view(np.uint32))


def test_parallel_chain_nonfinite_and_short():
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 31, 33, 255, 257, 4095, 4097):
        mat = _data("drift", max(n, 1), 1, rng)[:n]
        want, _ = _seq_partials(mat, 1, 1, 0, None)
        got, _ = _gpu_partials(mat, 1, 1, 0, None)
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), n
    mat = _data("drift", 30000, 2, rng)
    mat[12345, 0] = np.inf
    mat[23456, 1] = np.nan
    want, _ = _seq_partials(mat, 2, 2, 0, None)
    got, _ = _gpu_partials(mat, 2, 2, 0, None)
    assert np.array_equal(np.isnan(want), np.isnan(got))
    assert np.array_equal(want[~np.isnan(want)], got[~np.isnan(got)])


def test_parallel_chain_matches_the_sequential_kernel():
    rng = np.random.default_rng(9)
    mat = _data("walk", 500000, 1, rng)
    a, _ = _gpu_partials(mat, 1, 8, 0, None, impl=0)
    b, _ = _gpu_partials(mat, 1, 8, 0, None, impl=1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
