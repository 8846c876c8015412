"""Residual and ResidualSystematic resamplers, host halves (CPU only): the product's recurrences (csrc/host_resample.cpp,
called through the C ABI without a GPU) and the oracle's restatements against the UNMODIFIED reference classes
(novel_resampling.h:9-36, 76-104) compiled into oracle/_ref, with equally seeded generators."""
import ctypes as C

import numpy as np
import pytest

from oracle_lib import Oracle, Ref, ref_available
from tsdf_localization_b200 import capi

needs_ref = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference at build time)")


def weighted_cloud(n, kind, seed):
    rng = np.random.default_rng(seed)
    ps = rng.normal(size=(n, 7)).astype(np.float32)
    if kind == "flat":
        w = np.full(n, 1.0 / n)
    elif kind == "peaked":
        w = rng.random(n) ** 12
    elif kind == "sparse":
        w = np.where(rng.random(n) < 0.05, rng.random(n), 0.0)
        w[0] = 0.5
    else:
        w = rng.random(n)
    ps[:, 6] = (w / w.sum()).astype(np.float32)
    ps[:, 0] = np.arange(n, dtype=np.float32)        # x carries the index: parents can be read off the output
    return ps


def product_rs_counts(lib, w, u0):
    counts = np.empty(len(w), dtype=np.uint32)
    total = C.c_uint64(0)
    rc = lib.tsdfloc_residual_systematic_counts(w.ctypes.data_as(C.c_void_p), 1, len(w), C.c_float(u0), counts.ctypes.data_as(C.c_void_p),
                                                C.byref(total))
    assert rc == capi.OK
    assert int(counts.sum()) == total.value
    return counts


def product_residual_runs(lib, w, draws):
    n = len(w)
    it = iter(draws.tolist())
    cb = capi.INDEX_DRAW_FN(lambda _u: next(it))
    rp, rc_ = np.empty(n, dtype=np.uint32), np.empty(n, dtype=np.uint32)
    n_runs, n_draws = C.c_uint64(0), C.c_uint64(0)
    rc = lib.tsdfloc_residual_runs(w.ctypes.data_as(C.c_void_p), 1, n, cb, None, len(draws), rp.ctypes.data_as(C.c_void_p),
                                   rc_.ctypes.data_as(C.c_void_p), n, C.byref(n_runs), C.byref(n_draws))
    assert rc == capi.OK
    k = int(n_runs.value)
    return np.repeat(rp[:k], rc_[:k]), int(n_draws.value)


@needs_ref
@pytest.mark.parametrize("n", [1, 2, 7, 500, 4096, 65536])
@pytest.mark.parametrize("kind", ["flat", "peaked", "sparse", "uniform"])
def test_residual_systematic_matches_reference(n, kind):
    lib, oracle, ref = capi.load_library(), Oracle(), Ref()
    for seed in (1, 7):
        ps = weighted_cloud(n, kind, seed)
        m_ref, out_ref, u0 = ref.resample_method(2, ps, seed)
        parents_ref = out_ref[:, 0].astype(np.int64)
        assert np.array_equal(out_ref, ps[parents_ref])
        m_o, parents_o = oracle.residual_systematic_resample(ps[:, 6], u0)
        assert m_o == m_ref and np.array_equal(parents_o, parents_ref)
        counts = product_rs_counts(lib, np.ascontiguousarray(ps[:, 6]), u0)
        assert np.array_equal(np.repeat(np.arange(n), counts), parents_ref)


@needs_ref
@pytest.mark.parametrize("n", [1, 2, 7, 500, 4096, 65536])
@pytest.mark.parametrize("kind", ["flat", "peaked", "sparse", "uniform"])
def test_residual_matches_reference(n, kind):
    lib, oracle, ref = capi.load_library(), Oracle(), Ref()
    for seed in (3, 11):
        ps = weighted_cloud(n, kind, seed)
        m_ref, out_ref, _ = ref.resample_method(1, ps, seed)
        assert m_ref == n
        parents_ref = out_ref[:, 0].astype(np.int64)
        assert np.array_equal(out_ref, ps[parents_ref])
        draws = ref.uniform_index_draws(seed, n, 64 * n + 1024)
        m_o, parents_o, used_o = oracle.residual_resample(ps[:, 6], draws)
        assert m_o == n and np.array_equal(parents_o, parents_ref)
        parents_p, used_p = product_residual_runs(lib, np.ascontiguousarray(ps[:, 6]), draws)
        assert np.array_equal(parents_p, parents_ref) and used_p == used_o


def test_host_halves_reject_bad_input():
    lib = capi.load_library()
    w = np.array([0.5, -0.25, 0.75], dtype=np.float32)
    counts = np.empty(3, dtype=np.uint32)
    assert lib.tsdfloc_residual_systematic_counts(w.ctypes.data_as(C.c_void_p), 1, 3, C.c_float(0.3), counts.ctypes.data_as(C.c_void_p), None) == capi.E_BAD_ARG
    w = np.array([np.nan, 0.5], dtype=np.float32)
    assert lib.tsdfloc_residual_systematic_counts(w.ctypes.data_as(C.c_void_p), 1, 2, C.c_float(0.3), counts.ctypes.data_as(C.c_void_p), None) == capi.E_BAD_ARG
    # all-zero weights: the reference's Residual loop never terminates; the product gives up after max_draws
    z = np.zeros(4, dtype=np.float32)
    cb = capi.INDEX_DRAW_FN(lambda _u: 1)
    rp, rc_ = np.empty(4, dtype=np.uint32), np.empty(4, dtype=np.uint32)
    n_runs = C.c_uint64(0)
    assert lib.tsdfloc_residual_runs(z.ctypes.data_as(C.c_void_p), 1, 4, cb, None, 100, rp.ctypes.data_as(C.c_void_p),
                                     rc_.ctypes.data_as(C.c_void_p), 4, C.byref(n_runs), None) == capi.E_CAPACITY
    bad = capi.INDEX_DRAW_FN(lambda _u: 4)
    w = np.full(4, 0.25, dtype=np.float32)
    assert lib.tsdfloc_residual_runs(w.ctypes.data_as(C.c_void_p), 1, 4, bad, None, 100, rp.ctypes.data_as(C.c_void_p),
                                     rc_.ctypes.data_as(C.c_void_p), 4, C.byref(n_runs), None) == capi.E_BAD_ARG
