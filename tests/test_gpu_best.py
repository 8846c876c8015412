"""Arg-max ("best pose") of the normalised weights, mcl_3d.cpp:382-399: the first particle with the largest weight > 0."""
import numpy as np
import pytest

import common
from tsdf_localization_b200 import CudaEvaluator, SystematicResampler, synthetic as syn

pytestmark = pytest.mark.gpu


def reference_best(weights):
    """The reference loop, literally (mcl_3d.cpp:382-395)."""
    max_value, max_index = np.float32(0.0), -1
    for i, v in enumerate(weights):
        if v > max_value:
            max_value, max_index = v, i
    return max_index


def test_best_particle_after_evaluate():
    _, m = common.box_room(small=True)
    ev = CudaEvaluator(m)
    gt = (0.4, -0.3, 1.2, 0.01, -0.02, 0.4)
    pts, _ = syn.make_scan("vlp16", gt, room_lo=(-3.0, -2.5, 0.0), room_hi=(3.0, 2.5, 3.0), n_points=4096)
    for n in (1, 37, 5000):
        ps = syn.tracking_particles(n, gt, sigma_xy=0.2)
        ev.evaluate(ps, pts, syn.IDENTITY_TF)
        idx, pose, w = ev.best_particle()
        assert idx == reference_best(ps[:, 6])
        assert np.array_equal(pose, ps[idx, :6]) and w == ps[idx, 6]
    ev.close()


def test_best_particle_ties_and_zero_weights():
    """Resampler path (weights used as they are): ties resolve to the FIRST index; no positive weight -> -1."""
    _, m = common.box_room(small=True)
    ev = CudaEvaluator(m)
    rs = SystematicResampler(ev)
    rng = np.random.default_rng(2)
    n = 70000
    ps = rng.normal(size=(n, 7)).astype(np.float32)
    w = rng.random(n).astype(np.float32)
    w[[69999, 1024, 1023, 40000]] = np.float32(2.0)     # the same maximum in several scan tiles
    w /= w.sum(dtype=np.float64)
    ps[:, 6] = w.astype(np.float32)
    rs.resample(ps, u0=0.1 / n)
    idx, pose, bw = ev.best_particle()
    assert idx == 1023 == reference_best(ps[:, 6])
    assert np.array_equal(pose, ps[1023, :6]) and bw == ps[1023, 6]
    ev.close()
