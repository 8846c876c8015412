"""Motion update: the C oracle (oracle_motion_model + oracle_motion_apply) pinned against the
reference's own ParticleCloud::motionUpdate variants (src/particle_cloud.cpp:153-617, compiled verbatim into
oracle/_ref/libtsdf_ref_pc.so). The reference's draws are reproduced by equally seeded std::normal_distribution<> objects
(pc_harness.cpp), so the comparison is bit for bit. CPU only."""
import numpy as np
import pytest

from oracle_lib import Oracle, RefPC, ref_pc_path

pytestmark = pytest.mark.skipif(not ref_pc_path().exists(), reason="oracle/_ref/libtsdf_ref_pc.so not built (needs /root/reference)")

A_DEFAULT = [0.1] * 12
A_MIXED = [0.05, 0.3, 0.02, 0.11, 0.4, 0.07, 0.9, 0.13, 0.21, 0.6, 0.08, 0.33]

CASES = [
    (0, [0.8, 0.35]),                    # noise only: lin_scale, ang_scale
    (1, [1.3, -0.4]),                    # odometry: linear.x, angular.z
    (1, [0.0, 0.0]),                     # standing still: all sigmas 0 -> draws equal the means
    (2, [0.9, 0.25]),                    # IMU: linear_vel, angular_yaw
    (3, [0.7, 0.01, -0.02, 0.15]),       # noise + IMU deltas
]


def cloud(n, seed=0):
    rng = np.random.default_rng(seed)
    ps = np.zeros((n, 7), dtype=np.float32)
    ps[:, :3] = rng.uniform(-10, 10, size=(n, 3))
    ps[:, 3:5] = rng.normal(0, 0.05, size=(n, 2))
    ps[:, 5] = rng.uniform(-np.pi, np.pi, size=n)
    ps[:, 6] = rng.random(n)
    return ps


@pytest.mark.parametrize("variant,inputs", CASES)
@pytest.mark.parametrize("a", [A_DEFAULT, A_MIXED], ids=["a_default", "a_mixed"])
def test_motion_update_matches_reference(variant, inputs, a):
    ref, o = RefPC(), Oracle()
    ps = cloud(3000, seed=variant)
    dt = 0.1
    rp0 = np.array([0.3, -0.2, 0.0, 0.0, 0.0, 0.4], dtype=np.float32)
    want, rp_ref = ref.motion_update(variant, inputs, dt, a, 77, ps, rp0)
    # the reference computes FLOAT_T time_diff = (now - last).toSec() from the stub clock 100.0 -> 100.0 + dt
    time_diff = np.float32((100.0 + dt) - 100.0)
    mean, sigma, rp = o.motion_model(variant, inputs, time_diff, a, rp0)
    draws = ref.draws(77, mean, sigma, len(ps))
    got = o.motion_apply(ps, draws)
    assert got.tobytes() == want.tobytes()
    assert np.array_equal(got[:, 6], ps[:, 6])            # weights ride along (:604)
    if variant in (1, 2):
        assert rp.tobytes() == rp_ref.tobytes()
    else:
        assert rp_ref.tobytes() == rp0.tobytes()          # the noise-only variants leave the reference pose alone


def test_gimbal_lock_branch():
    """getAngleFromMat's |mat[8]| >= 1 branch (util.cpp:88-103): pitch = +-pi/2."""
    o = Oracle()
    ps = np.zeros((2, 7), dtype=np.float32)
    ps[0, 4] = np.float32(np.pi / 2)
    ps[1, 4] = np.float32(-np.pi / 2)
    out = o.motion_apply(ps, np.zeros((2, 6)))
    assert abs(abs(out[0, 4]) - np.pi / 2) < 1e-3 and abs(abs(out[1, 4]) - np.pi / 2) < 1e-3
