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


# ---- Wheel / Metropolis / Rejection (the rest of mcl_3d's resampling_method switch, src/mcl_3d.cpp:243-263) -----------------

def product_drawn_parents(lib, method, w, draws, steps=50, max_draws=0):
    n = len(w)
    parents = np.empty(n, dtype=np.uint32)
    real = C.cast(draws.real_wheel_ptr if method == capi.RESAMPLE_WHEEL else draws.real_ptr, capi.REAL_DRAW_FN)
    index = C.cast(draws.index_ptr, capi.INDEX_DRAW_FN)
    wp, pp = w.ctypes.data_as(C.c_void_p), parents.ctypes.data_as(C.c_void_p)
    if method == capi.RESAMPLE_WHEEL:
        rc = lib.tsdfloc_wheel_parents(wp, 1, n, real, draws.handle, pp)
    elif method == capi.RESAMPLE_METROPOLIS:
        rc = lib.tsdfloc_metropolis_parents(wp, 1, n, steps, real, index, draws.handle, pp)
    else:
        rc = lib.tsdfloc_rejection_parents(wp, 1, n, real, index, draws.handle, max_draws, pp, None)
    assert rc == capi.OK
    return parents


DRAWN = [(capi.RESAMPLE_WHEEL, "wheel"), (capi.RESAMPLE_METROPOLIS, "metropolis"), (capi.RESAMPLE_REJECTION, "rejection")]


@needs_ref
@pytest.mark.parametrize("n", [1, 2, 7, 500, 4096])
@pytest.mark.parametrize("kind", ["flat", "peaked", "sparse", "uniform"])
@pytest.mark.parametrize("method,name", DRAWN)
def test_drawn_resamplers_match_reference(method, name, n, kind):
    """The verbatim WheelResampler / MetropolisResampler(steps) / RejectionResampler with a seeded std::mt19937 against the
    oracle's restatements and the product's host halves fed the same generator through the same distribution objects: parents
    identical, and exactly as many draws consumed."""
    lib, oracle, ref = capi.load_library(), Oracle(), Ref()
    if name == "rejection" and kind in ("peaked", "sparse") and n > 500:
        pytest.skip("O(n * sup_w / mean_w) draws: minutes in the reference itself")
    steps = 50 if n <= 500 else 7
    ref.set_metropolis_steps(steps)
    for seed in (5, 23):
        ps = weighted_cloud(n, kind, seed)
        if name == "wheel":
            ps[:, 6] *= np.float32(0.9)        # some draws exceed the last running sum: those slots keep their own particle
        m_ref, out_ref, _ = ref.resample_method(method, ps, seed)
        assert m_ref == n
        parents_ref = out_ref[:, 0].astype(np.int64)
        assert np.array_equal(out_ref, ps[parents_ref])
        w = np.ascontiguousarray(ps[:, 6])
        d_o, d_p = ref.draws(seed, n), ref.draws(seed, n)
        parents_o = oracle.drawn_resample(method, w, d_o, steps)
        assert np.array_equal(parents_o, parents_ref)
        parents_p = product_drawn_parents(lib, method, w, d_p, steps)
        assert np.array_equal(parents_p, parents_ref)
        assert d_p.used() == d_o.used()
    ref.set_metropolis_steps(50)


@pytest.mark.parametrize("n", [1, 3, 1000, 20000])
@pytest.mark.parametrize("method,name", DRAWN)
def test_drawn_host_halves_match_oracle(method, name, n):
    """The same comparison on the oracle's own draw source (no reference build needed), incl. weight vectors the wheel's O(n)
    search must not get wrong: sums far below / above 1, zero runs, negative weights (non-monotone running sums)."""
    lib, oracle = capi.load_library(), Oracle()
    rng = np.random.default_rng(n)
    cases = {"normalised": rng.random(n), "short": rng.random(n) * 0.3, "long": rng.random(n) * 4.0,
             "zero runs": np.where(rng.random(n) < 0.7, 0.0, rng.random(n)) + np.eye(1, n, n - 1).ravel(), "one": np.eye(1, n, n // 2).ravel()}
    if name == "wheel":
        cases["signed"] = rng.normal(size=n) + 0.2
    for label, w in cases.items():
        if name == "rejection" and label == "one" and n > 1000:
            continue
        scale = {"short": 0.3, "long": 4.0}.get(label, 1.0)
        w32 = (w / w.sum() * scale).astype(np.float32) if label != "signed" else (w / n).astype(np.float32)
        d_o, d_p = oracle.draws(n + 1, n), oracle.draws(n + 1, n)
        parents_o = oracle.drawn_resample(method, w32, d_o, 9)
        parents_p = product_drawn_parents(lib, method, w32, d_p, 9)
        assert np.array_equal(parents_p, parents_o), label
        assert d_p.used() == d_o.used(), label


def test_drawn_host_halves_reject_bad_input():
    lib, oracle = capi.load_library(), Oracle()
    w = np.full(4, 0.25, dtype=np.float32)
    parents = np.empty(4, dtype=np.uint32)
    wp, pp = w.ctypes.data_as(C.c_void_p), parents.ctypes.data_as(C.c_void_p)
    half = capi.REAL_DRAW_FN(lambda _u: 0.5)
    bad = capi.INDEX_DRAW_FN(lambda _u: 4)
    none_r, none_i = C.cast(None, capi.REAL_DRAW_FN), C.cast(None, capi.INDEX_DRAW_FN)
    assert lib.tsdfloc_wheel_parents(wp, 1, 4, none_r, None, pp) == capi.E_BAD_ARG
    assert lib.tsdfloc_wheel_parents(wp, 1, 0, half, None, pp) == capi.E_BAD_ARG
    assert lib.tsdfloc_metropolis_parents(wp, 1, 4, 3, half, bad, None, pp) == capi.E_BAD_ARG
    assert lib.tsdfloc_metropolis_parents(wp, 1, 4, 3, half, none_i, None, pp) == capi.E_BAD_ARG
    # all-negative weights: sup_w stays 0, w / 0 = -inf, the reference's loop never ends; max_draws bounds it
    neg = np.full(4, -0.25, dtype=np.float32)
    zero = capi.INDEX_DRAW_FN(lambda _u: 0)
    used = C.c_uint64(0)
    assert lib.tsdfloc_rejection_parents(neg.ctypes.data_as(C.c_void_p), 1, 4, half, zero, None, 100, pp, C.byref(used)) == capi.E_CAPACITY
    assert used.value == 100
    # all-zero weights: 0 / 0 is NaN, `u > NaN` is false, every slot keeps itself — without a single index draw
    z = np.zeros(4, dtype=np.float32)
    assert lib.tsdfloc_rejection_parents(z.ctypes.data_as(C.c_void_p), 1, 4, half, bad, None, 0, pp, C.byref(used)) == capi.OK
    assert used.value == 0 and np.array_equal(parents, np.arange(4))
    # Metropolis against a zero-weight particle 0: w / 0 = inf accepts every proposal, 0 / 0 = NaN none
    w0 = np.array([0.0, 0.5, 0.0, 0.5], dtype=np.float32)
    seq = iter([1, 2, 3, 2] * 4)
    it = capi.INDEX_DRAW_FN(lambda _u: next(seq))
    assert lib.tsdfloc_metropolis_parents(w0.ctypes.data_as(C.c_void_p), 1, 4, 4, half, it, None, pp) == capi.OK
    assert np.array_equal(parents, [3, 3, 3, 3])


def test_wheel_search_equals_linear_walk_on_hostile_floats():
    """The wheel's O(n) search (running maximum + guide table + bracketed search) against the oracle's plain walk on weights and
    draws no sane caller produces: negative, zero, subnormal, huge, infinite and NaN weights; draws of 0, 1, > 1, < 0, NaN."""
    lib, oracle = capi.load_library(), Oracle()
    rng = np.random.default_rng(99)
    specials = np.array([0.0, -0.0, 1e-45, -1e-45, 1e-38, 3.4e38, -3.4e38, np.inf, -np.inf, np.nan, 1.0, 0.5, -0.5, 2.0 ** -24], dtype=np.float32)
    draw_specials = np.array([0.0, -0.0, 1.0, np.nextafter(np.float32(1), np.float32(0)), 1.5, -0.25, np.nan, np.inf, 1e-45, 2.0 ** -24], dtype=np.float32)
    for trial in range(400):
        n = int(rng.integers(1, 40))
        w = (rng.normal(size=n) * rng.choice([1e-3, 0.1, 1.0])).astype(np.float32)
        if trial % 2:
            w = np.abs(w)
        k = rng.integers(0, n, size=rng.integers(0, 4))
        w[k] = rng.choice(specials, size=len(k))
        u = rng.random(n).astype(np.float32)
        j = rng.integers(0, n, size=rng.integers(0, 3))
        u[j] = rng.choice(draw_specials, size=len(j))
        it_o, it_p = iter(u.tolist()), iter(u.tolist())
        cb_o, cb_p = capi.REAL_DRAW_FN(lambda _u: next(it_o)), capi.REAL_DRAW_FN(lambda _u: next(it_p))
        want = np.empty(n, dtype=np.uint32)
        oracle.lib.oracle_wheel_resample(w.ctypes.data_as(C.c_void_p), n, C.cast(cb_o, C.c_void_p), None, want.ctypes.data_as(C.c_void_p))
        got = np.empty(n, dtype=np.uint32)
        assert lib.tsdfloc_wheel_parents(w.ctypes.data_as(C.c_void_p), 1, n, cb_p, None, got.ctypes.data_as(C.c_void_p)) == capi.OK
        assert np.array_equal(got, want), (trial, w, u, got, want)
