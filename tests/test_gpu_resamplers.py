"""Residual / ResidualSystematic / Wheel / Metropolis / Rejection resampling end to end on the GPU (host half over the weights +
k_expand_runs),
through the Python mirror of the reference classes, against the oracle (pinned to the verbatim reference classes in
tests/test_resamplers_host.py): parents identical, copies are byte copies of the parents."""
import ctypes as C

import numpy as np
import pytest

import common
from tsdf_localization_b200 import (CudaEvaluator, MetropolisResampler, RejectionResampler, ResidualResampler, ResidualSystematicResampler,
                                    WheelResampler, capi, synthetic as syn)
from test_resamplers_host import weighted_cloud

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ev():
    _, m = common.box_room(small=True)
    e = CudaEvaluator(m)
    yield e
    e.close()


@pytest.mark.parametrize("n", [1, 2, 500, 4097, 262144])
@pytest.mark.parametrize("kind", ["flat", "peaked", "sparse"])
def test_residual_systematic_weighted_cloud(oracle, ev, n, kind):
    ps = weighted_cloud(n, kind, n)
    rs = ResidualSystematicResampler(ev)
    for u0 in (0.0, 0.37, float(np.nextafter(np.float32(1.0), np.float32(0.0)))):
        out, parents = rs.resample(ps, u0=u0, want_parents=True)
        m_ref, parents_ref = oracle.residual_systematic_resample(ps[:, 6], u0)
        assert len(out) == m_ref and np.array_equal(parents, parents_ref)
        assert np.array_equal(out, ps[parents_ref])


@pytest.mark.parametrize("n", [1, 2, 500, 4097, 262144])
@pytest.mark.parametrize("kind", ["flat", "peaked", "sparse"])
def test_residual_weighted_cloud(oracle, ev, n, kind):
    ps = weighted_cloud(n, kind, n + 1)
    draws = np.random.default_rng(n).integers(0, n, size=64 * n + 1024).astype(np.uint64)
    out, parents = ResidualResampler(ev).resample(ps, index_draws=draws, want_parents=True)
    m_ref, parents_ref, _ = oracle.residual_resample(ps[:, 6], draws)
    assert m_ref == n and len(out) == n and np.array_equal(parents, parents_ref)
    assert np.array_equal(out, ps[parents_ref])


DRAWN = [(3, WheelResampler), (4, MetropolisResampler), (5, RejectionResampler)]


@pytest.mark.parametrize("n", [1, 2, 500, 4097, 65536])
@pytest.mark.parametrize("kind", ["flat", "uniform", "sparse"])
@pytest.mark.parametrize("method,cls", DRAWN, ids=["wheel", "metropolis", "rejection"])
def test_drawn_resamplers_weighted_cloud(oracle, ev, method, cls, n, kind):
    """Wheel / Metropolis / Rejection end to end (host half over the weights + k_expand_runs) against the oracle's restatements
    (pinned to the verbatim reference classes in tests/test_resamplers_host.py) on equally seeded draw sources."""
    if method != 4 and n > 4097:
        n = 16384          # the oracle walks the wheel in O(n^2) like the reference; sparse rejection needs ~20 draws per slot
    ps = weighted_cloud(n, kind, n + method)
    steps = 50 if n <= 500 else 5
    rs = cls(ev, steps) if method == 4 else cls(ev)
    d_o, d_p = oracle.draws(n + method, n), oracle.draws(n + method, n)
    parents_ref = oracle.drawn_resample(method, ps[:, 6], d_o, steps)
    out, parents = rs.resample(ps, draws=d_p.source(), want_parents=True)
    assert len(out) == n and np.array_equal(parents, parents_ref)
    assert np.array_equal(out, ps[parents_ref])
    assert d_p.used() == d_o.used()


def test_drawn_resamplers_on_the_resident_set(oracle):
    """The node's flow with resampling_method 0 / 4 / 5: sensor update, then the resampler on the set left on the device."""
    _, m = common.box_room()
    e = CudaEvaluator(m)
    ps, pts, _ = common.config_c1()
    mine = ps.copy()
    e.evaluate(mine, pts, syn.IDENTITY_TF)
    n = len(mine)
    for method, cls in DRAWN:
        rs = cls(e)
        d_o, d_p = oracle.draws(method, n), oracle.draws(method, n)
        parents_ref = oracle.drawn_resample(method, mine[:, 6], d_o, 50)
        out, parents = rs.resample_resident(n, draws=d_p.source(), want_parents=True)
        assert len(out) == n and np.array_equal(parents, parents_ref) and np.array_equal(out, mine[parents_ref])
    # default draws (a seeded numpy Generator through Python callbacks): a valid resampled set, reproducible by seed
    a = WheelResampler(e, seed=3).resample_resident(n)
    b = WheelResampler(e, seed=3).resample_resident(n)
    assert np.array_equal(a, b) and len(a) == n
    assert (a[:, None, :] == mine[None, :, :]).all(-1).any(-1).all()        # every output is one of the particles
    e.close()


def test_drawn_resampler_errors(ev):
    ps = weighted_cloud(64, "flat", 0)
    ps[:, 6] = -1.0                                   # the reference's rejection loop never ends on all-negative weights
    with pytest.raises(capi.TsdflocError) as ei:
        RejectionResampler(ev, seed=1).resample(ps, max_draws=1000)
    assert ei.value.status == capi.E_CAPACITY
    lib = ev._lib
    n_out = C.c_uint64(0)
    out = np.empty((64, 7), dtype=np.float32)
    rc = lib.tsdfloc_resample_drawn(ev.ctx, capi.RESAMPLE_RESIDUAL, ps.ctypes.data_as(C.c_void_p), 64, C.byref(capi.Draws()),
                                    out.ctypes.data_as(C.c_void_p), 64, C.byref(n_out), None)
    assert rc == capi.E_BAD_ARG


def test_resident_set_after_sensor_update(oracle):
    """The flow of the node: sensor update, then the configured resampler on the particle set left on the device — only
    4 B per particle visit the host."""
    _, m = common.box_room()
    e = CudaEvaluator(m)
    ps, pts, _ = common.config_c1()
    mine = ps.copy()
    e.evaluate(mine, pts, syn.IDENTITY_TF)
    n = len(mine)
    out, parents = ResidualSystematicResampler(e).resample_resident(n, u0=0.61, want_parents=True)
    m_ref, parents_ref = oracle.residual_systematic_resample(mine[:, 6], 0.61)
    assert len(out) == m_ref and np.array_equal(parents, parents_ref) and np.array_equal(out, mine[parents_ref])
    draws = np.random.default_rng(0).integers(0, n, size=64 * n).astype(np.uint64)
    out, parents = ResidualResampler(e).resample_resident(n, index_draws=draws, want_parents=True)
    m_ref, parents_ref, _ = oracle.residual_resample(mine[:, 6], draws)
    assert len(out) == m_ref == n and np.array_equal(parents, parents_ref) and np.array_equal(out, mine[parents_ref])
    e.close()


def test_errors(ev):
    ps = weighted_cloud(64, "flat", 0)
    ps[:, 6] = 0.0
    with pytest.raises(capi.TsdflocError) as ei:
        ResidualResampler(ev).resample(ps, index_draws=np.zeros(10000, dtype=np.uint64))
    assert ei.value.status in (capi.E_NO_VALID_PARTICLE, capi.E_BAD_ARG)
    ps[:, 6] = -1.0
    with pytest.raises(capi.TsdflocError):
        ResidualSystematicResampler(ev).resample(ps, u0=0.5)
