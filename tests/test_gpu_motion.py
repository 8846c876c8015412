"""Motion update on the GPU (k_motion_apply through tsdfloc_motion_update): parity mode (injected draws) against the oracle
and, where oracle/_ref/libtsdf_ref_pc.so travelled along, against the reference's own ParticleCloud::motionUpdate; Philox mode
through its statistics and determinism.

Tolerance: positions and weights bit-exact; Euler angles bit-exact except where CUDA's fp64 asin/atan2/sincos (<= 2 ulp in
double) and glibc's round to different fp32 neighbours — at most 1 fp32 ulp, on at most 1e-4 of the particles."""
import numpy as np
import pytest

import common
from oracle_lib import RefPC, ref_pc_path
from test_motion_oracle import A_DEFAULT, A_MIXED, CASES, cloud
from tsdf_localization_b200 import CudaEvaluator, ParticleCloud, capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def evaluator():
    _, m = common.box_room(small=True)
    ev = CudaEvaluator(m)
    yield ev
    ev.close()


def assert_poses_match(got, want):
    assert np.array_equal(got[:, 6], want[:, 6])
    assert got[:, :3].tobytes() == want[:, :3].tobytes(), "positions differ"
    diff = got[:, 3:6] != want[:, 3:6]
    if diff.any():
        ulp = np.abs(got[:, 3:6].view(np.int32).astype(np.int64) - want[:, 3:6].view(np.int32).astype(np.int64))
        assert ulp.max() <= 1, f"angle off by {ulp.max()} ulp"
        assert diff.any(axis=1).mean() <= 1e-4, f"{diff.any(axis=1).sum()} of {len(got)} particles differ in the last bit"


def run_variant(pc, variant, inputs, dt, draws):
    if variant == capi.MOTION_NOISE:
        pc.motionUpdateNoise(*inputs, time_diff=dt, draws=draws)
    elif variant == capi.MOTION_ODOM:
        pc.motionUpdateOdom(*inputs, time_diff=dt, draws=draws)
    elif variant == capi.MOTION_IMU:
        pc.motionUpdateImu(*inputs, time_diff=dt, draws=draws)
    else:
        pc.motionUpdateNoiseImu(*inputs, time_diff=dt, draws=draws)


@pytest.mark.parametrize("variant,inputs", CASES)
def test_injected_draws_match_oracle(oracle, evaluator, variant, inputs):
    ps = cloud(20000, seed=10 + variant)
    rng = np.random.default_rng(variant)
    mean, sigma, _ = oracle.motion_model(variant, inputs, np.float32(0.1), A_MIXED)
    draws = mean + sigma * rng.normal(size=(len(ps), 6))
    want = oracle.motion_apply(ps, draws)
    pc = ParticleCloud(evaluator, ps.copy())
    pc.setAParams(*A_MIXED)
    run_variant(pc, variant, inputs, 0.1, draws)
    assert_poses_match(pc.particles(), want)


@pytest.mark.skipif(not ref_pc_path().exists(), reason="oracle/_ref/libtsdf_ref_pc.so not built")
@pytest.mark.parametrize("variant,inputs", CASES)
def test_matches_reference_particle_cloud(oracle, evaluator, variant, inputs):
    """The reference's own motionUpdate (verbatim, seeded mt19937) vs the GPU fed with the same draws."""
    ref = RefPC()
    ps = cloud(5000, seed=20 + variant)
    rp0 = np.array([0.3, -0.2, 0.0, 0.0, 0.0, 0.4], dtype=np.float32)
    want, rp_ref = ref.motion_update(variant, inputs, 0.1, A_DEFAULT, 5, ps, rp0)
    pc = ParticleCloud(evaluator, ps.copy())
    pc.ref_pose[:] = rp0
    time_diff = float(np.float32((100.0 + 0.1) - 100.0))
    # the same (mean, sigma) the product computes feed the reference-identical generator
    probe = ParticleCloud(evaluator)
    probe.ref_pose[:] = rp0
    mean, sigma = probe.model(variant, inputs, time_diff)
    run_variant(pc, variant, inputs, time_diff, ref.draws(5, mean, sigma, len(ps)))
    assert_poses_match(pc.particles(), want)
    assert pc.ref_pose.tobytes() == rp_ref.tobytes()


def test_gimbal_lock_and_edge_sizes(oracle, evaluator):
    ps = np.zeros((3, 7), dtype=np.float32)
    ps[0, 4] = np.float32(np.pi / 2)
    ps[1, 4] = np.float32(-np.pi / 2)
    ps[2, :6] = (1, 2, 3, 0.1, 0.2, 0.3)
    draws = np.zeros((3, 6))
    draws[2] = (0.5, -0.25, 0.125, 0.01, -0.02, 0.03)
    want = oracle.motion_apply(ps, draws)
    pc = ParticleCloud(evaluator, ps.copy())
    pc.motionUpdateNoise(0.0, 0.0, time_diff=0.1, draws=draws)
    assert_poses_match(pc.particles(), want)
    one = ps[2:3].copy()
    pc = ParticleCloud(evaluator, one)
    pc.motionUpdateNoise(0.0, 0.0, time_diff=0.1)          # sigma == 0: every sample equals its mean (0) -> pose unchanged
    assert np.allclose(one[0, :6], ps[2, :6], atol=1e-6)


def test_philox_statistics_and_determinism(evaluator):
    n = 200000
    base = np.zeros((n, 7), dtype=np.float32)
    base[:, 6] = 1.0 / n
    a = ParticleCloud(evaluator, base.copy(), seed=123)
    a.setAParams(*A_MIXED)
    a.motionUpdateOdom(1.3, -0.4, time_diff=0.5)
    probe = ParticleCloud(evaluator)
    probe.setAParams(*A_MIXED)
    mean, sigma = probe.model(capi.MOTION_ODOM, [1.3, -0.4], 0.5)
    out = a.particles()
    # particles start at the origin with zero rotation: the displacement IS the (dx, dy, dz) sample, the angles the rotation sample
    for k in range(3):
        assert abs(out[:, k].mean() - mean[k]) < 5 * sigma[k] / np.sqrt(n) + 1e-6
        assert abs(out[:, k].std() - sigma[k]) < 0.02 * sigma[k] + 1e-6
    assert abs(out[:, 5].mean() - mean[5]) < 5 * sigma[5] / np.sqrt(n) + 1e-4
    assert abs(np.corrcoef(out[:, 0], out[:, 1])[0, 1]) < 0.02
    # keyed by (seed, sequence, particle index): same key -> same samples, independent of the launch size
    b = ParticleCloud(evaluator, base[:1000].copy(), seed=123)
    b.setAParams(*A_MIXED)
    b.motionUpdateOdom(1.3, -0.4, time_diff=0.5)
    assert b.particles().tobytes() == out[:1000].tobytes()
    c = ParticleCloud(evaluator, base[:1000].copy(), seed=124)
    c.setAParams(*A_MIXED)
    c.motionUpdateOdom(1.3, -0.4, time_diff=0.5)
    assert c.particles().tobytes() != out[:1000].tobytes()
    b.m_particles = base[:1000].copy()
    b.motionUpdateOdom(1.3, -0.4, time_diff=0.5)           # second update: next sequence number -> fresh samples
    assert b.particles().tobytes() != out[:1000].tobytes()


def test_initialize_three_modes(evaluator):
    """ParticleCloud::initialize x 3 (particle_cloud.cpp:32-148): weights 1/n, sample statistics, determinism."""
    n = 100000
    centre = (1.0, -2.0, 0.5, 0.01, -0.02, 0.7)
    pc = ParticleCloud(evaluator, seed=9)
    sig = (0.5, 0.25, 0.1, 0.02, 0.03, 0.4)
    ps = pc.initialize(n, centre, sig, capi.INIT_NORMAL)
    assert ps.shape == (n, 7) and (ps[:, 6] == np.float32(1.0 / n)).all() and not pc.ref_pose.any()
    for k in range(6):
        assert abs(ps[:, k].mean() - centre[k]) < 5 * sig[k] / np.sqrt(n)
        assert abs(ps[:, k].std() - sig[k]) < 0.02 * sig[k]
    half = (2.0, 3.0, 0.5, 0.1, 0.2, np.pi)
    ps = pc.initialize(n, centre, half, capi.INIT_UNIFORM)
    for k in range(6):
        assert ps[:, k].min() >= centre[k] - half[k] - 1e-5 and ps[:, k].max() <= centre[k] + half[k] + 1e-5
        assert abs(ps[:, k].std() - half[k] / np.sqrt(3)) < 0.02 * half[k]
    rng = np.random.default_rng(0)
    free = rng.uniform(-5, 5, size=(777, 3)).astype(np.float32)
    ps = pc.initialize(n, centre, half, capi.INIT_FREE_MAP, free_map=free)
    want = free.copy()
    want[:, 2] = (free[:, 2].astype(np.float64) - 0.5).astype(np.float32)
    rows = {r.tobytes() for r in want}
    assert all(p.tobytes() in rows for p in ps[:2000, :3])
    counts = np.unique(ps[:, 0], return_counts=True)[1]
    assert len(counts) == len(np.unique(free[:, 0])) and counts.min() > 0.5 * n / len(free)      # every voxel drawn, roughly evenly
    again = ParticleCloud(evaluator, seed=9)
    again.sequence = pc.sequence - 1
    assert again.initialize(n, centre, half, capi.INIT_FREE_MAP, free_map=free).tobytes() == ps.tobytes()
    with pytest.raises(capi.TsdflocError):
        pc.initialize(10, centre, half, capi.INIT_FREE_MAP)
