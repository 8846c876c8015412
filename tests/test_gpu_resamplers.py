"""Residual / ResidualSystematic resampling end to end on the GPU (host recurrence over the weights + k_expand_runs),
through the Python mirror of the reference classes, against the oracle (pinned to the verbatim reference classes in
tests/test_resamplers_host.py): parents identical, copies are byte copies of the parents."""
import numpy as np
import pytest

import common
from tsdf_localization_b200 import CudaEvaluator, ResidualResampler, ResidualSystematicResampler, capi, synthetic as syn
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
