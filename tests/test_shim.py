"""The drop-in seam itself: the reference's OWN TSDFEvaluator / Resampler classes (compiled unmodified from /root/reference
into oracle/_ref/libtsdf_ref_shim.so) linked against the product's C++ shim (tsdf_localization_b200/shim) + libtsdfloc.so.
`TSDFEvaluator::evaluate(particles, points, tf, use_cuda=true)` (tsdf_evaluator.cpp:82) then runs on the B200, and its result
is compared with the same object's CPU/OpenMP branch (use_cuda=false) — the reference checking its replacement."""
import numpy as np
import pytest

import common
from oracle_lib import Ref, ref_shim_path
from tsdf_localization_b200 import synthetic as syn

pytestmark = pytest.mark.skipif(not ref_shim_path().exists(), reason="oracle/_ref/libtsdf_ref_shim.so not built")


def _has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.fixture(scope="module")
def shim():
    return Ref(shim=True)


@pytest.fixture(scope="module")
def small_map(shim):
    spec, _ = common.box_room(small=True)
    rm = shim.map_create(spec.min, spec.max, spec.resolution, spec.init_value)
    assert shim.map_set_data(rm, spec.cells) == 0
    yield rm
    shim.map_destroy(rm)


def _workload(n=512, p=3000):
    """Tight tracking cloud: no lookup lands at a negative axis offset (the small room's box is only 0.7 m wider than the room
    in y), so the reference's CPU branch is well defined everywhere and must agree with the GPU branch."""
    gt = (0.4, -0.3, 1.2, 0.01, -0.02, 0.4)
    pts, _ = syn.make_scan("vlp16", gt, room_lo=(-3.0, -2.5, 0.0), room_hi=(3.0, 2.5, 3.0), n_points=p)
    ps = syn.tracking_particles(n, gt, sigma_xy=0.05, sigma_z=0.05, sigma_yaw=0.03)
    return ps, pts


def test_constructor_fails_loudly_without_gpu(shim, small_map):
    if _has_gpu():
        pytest.skip("a GPU is present")
    ev = shim.eval_create(small_map)     # TSDFEvaluator's ctor constructs CudaEvaluator unconditionally (tsdf_evaluator.h:78)
    assert not ev
    err = shim.last_error()
    assert "Error while creating the CUDA context for the map!" in err and "no CPU fallback" in err


@pytest.mark.gpu
def test_reference_facade_gpu_branch_matches_its_cpu_branch(shim, small_map):
    ev = shim.eval_create(small_map)
    assert ev, shim.last_error()
    ps, pts = _workload()
    rc_c, cpu, pose_c, err_c = shim.evaluate(ev, ps, pts, syn.CALIB_TF, use_cuda=False)
    rc_g, gpu, pose_g, err_g = shim.evaluate(ev, ps, pts, syn.CALIB_TF, use_cuda=True)
    assert rc_c == 0 and rc_g == 0, (err_c, err_g)
    assert np.array_equal(gpu[:, :6], ps[:, :6])
    rel = common.rel_err(gpu[:, 6], cpu[:, 6])
    assert rel.max() <= 1e-5, f"normalised weights: max rel err {rel.max():.2e}"      # north-star tolerance
    np.testing.assert_allclose(pose_g[3:], pose_c[3:], atol=1e-5)       # orientation (race-free in the reference)
    # position: the reference's CPU xyz mean is racy (lost updates, see test_oracle_vs_ref); check against the weights instead
    xyz = (gpu[:, :3].astype(np.float64) * gpu[:, 6:7].astype(np.float64)).sum(0)
    np.testing.assert_allclose(pose_g[:3], xyz, atol=1e-5)
    shim.eval_destroy(ev)


@pytest.mark.gpu
def test_reference_facade_gpu_branch_errors(shim, small_map):
    ev = shim.eval_create(small_map, 0.9, 0.0, 0.0, 100.0)
    ps, pts = _workload(16, 64)
    far = ps.copy()
    far[:, :3] += 500.0
    rc, _, _, err = shim.evaluate(ev, far, pts, syn.IDENTITY_TF, use_cuda=True)
    assert rc == 1 and err == "No particle is valid!"
    # empty scan: default pose, weights untouched (cuda_evaluator.cu:122-125)
    ps[:, 6] = 0.25
    rc, out, pose, _ = shim.evaluate(ev, ps, np.zeros((0, 3), dtype=np.float32), syn.IDENTITY_TF, use_cuda=True)
    assert rc == 0 and np.array_equal(out, ps) and not pose.any()
    shim.eval_destroy(ev)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [500, 4096, 100000])
def test_gpu_resampler_subclass_matches_reference_resampler(shim, small_map, n):
    ev = shim.eval_create(small_map)      # owns the GPU context the resampler borrows
    rng = np.random.default_rng(n)
    ps = np.zeros((n, 7), dtype=np.float32)
    ps[:, :6] = rng.normal(size=(n, 6))
    w = rng.exponential(size=n) ** 2
    ps[:, 6] = (w / w.sum()).astype(np.float32)
    for seed in (1, 7):
        m_ref, out_ref, _ = shim.systematic_resample(ps, seed)
        m_gpu, out_gpu = shim.gpu_systematic_resample(ps, seed)
        assert m_gpu == m_ref
        assert np.array_equal(out_gpu, out_ref)
    shim.eval_destroy(ev)


@pytest.mark.gpu
@pytest.mark.parametrize("method", [1, 2], ids=["residual", "residual_systematic"])
@pytest.mark.parametrize("n", [500, 4096, 100000])
def test_gpu_residual_resamplers_match_reference_classes(shim, small_map, n, method):
    """GpuResidualResampler / GpuResidualSystematicResampler (shim/tsdfloc_shim.h: host recurrence + device expansion) through
    the reference's Resampler interface against the reference's own ResidualResampler / ResidualSystematicResampler
    (novel_resampling.h:9-36, 76-104 — the methods a default-configured node selects) with equal seeds: identical outputs."""
    ev = shim.eval_create(small_map)
    rng = np.random.default_rng(n + method)
    ps = np.zeros((n, 7), dtype=np.float32)
    ps[:, :6] = rng.normal(size=(n, 6))
    w = rng.exponential(size=n) ** 2
    ps[:, 6] = (w / w.sum()).astype(np.float32)
    for seed in (1, 7):
        m_ref, out_ref, _ = shim.resample_method(method, ps, seed)
        m_gpu, out_gpu = shim.gpu_resample_method(method, ps, seed)
        assert m_gpu == m_ref
        assert np.array_equal(out_gpu, out_ref)
    shim.eval_destroy(ev)


@pytest.mark.gpu
@pytest.mark.parametrize("method", [3, 4, 5], ids=["wheel", "metropolis", "rejection"])
@pytest.mark.parametrize("n", [500, 4096, 20000])
def test_gpu_drawn_resamplers_match_reference_classes(shim, small_map, n, method):
    """GpuWheelResampler / GpuMetropolisResampler(steps) / GpuRejectionResampler (shim/tsdfloc_shim.h) through the reference's
    Resampler interface against the reference's own WheelResampler (src/resampling/wheel_resampler.cpp), MetropolisResampler and
    RejectionResampler (novel_resampling.h:106-189) — cases 0, 4 and default of src/mcl_3d.cpp:243-263 — with equal seeds:
    identical outputs. (20,000 particles bound the reference's own O(n^2) wheel walk.)"""
    ev = shim.eval_create(small_map)
    rng = np.random.default_rng(n + method)
    ps = np.zeros((n, 7), dtype=np.float32)
    ps[:, :6] = rng.normal(size=(n, 6))
    w = rng.exponential(size=n)
    ps[:, 6] = (w / w.sum()).astype(np.float32)
    shim.set_metropolis_steps(50 if n <= 4096 else 10)
    for seed in (1, 7):
        m_ref, out_ref, _ = shim.resample_method(method, ps, seed)
        m_gpu, out_gpu = shim.gpu_resample_method(method, ps, seed)
        assert m_gpu == m_ref == n
        assert np.array_equal(out_gpu, out_ref)
    shim.set_metropolis_steps(50)
    shim.eval_destroy(ev)


@pytest.mark.gpu
@pytest.mark.parametrize("near", [0, 400])
def test_reference_evaluateParticles_gpu_branch_matches_its_cpu_branch(shim, small_map, near):
    """TSDFEvaluatorB200::evaluateParticles (GPU scan reduction + evaluation, the reduced scan never leaves the device)
    against the reference's own evaluateParticles CPU branch (tsdf_evaluator.cpp:247-378) on the same facade object.
    near > 0: points closer than 1 m desynchronise the reference's ring iterator; RING_DESYNC mode reproduces that."""
    from test_reduce_oracle import scan_with_rings
    from oracle_lib import Oracle
    ev = shim.eval_create_cell(small_map, 0.064)
    assert ev, shim.last_error()
    gt = (0.4, -0.3, 1.2, 0.01, -0.02, 0.4)
    pts, ring = syn.make_scan("vlp16", gt, room_lo=(-3.0, -2.5, 0.0), room_hi=(3.0, 2.5, 3.0))
    if near:
        rng = np.random.default_rng(near)
        d = rng.normal(size=(near, 3)).astype(np.float32)
        d *= (rng.uniform(0.05, 0.99, size=(near, 1)) / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        pos = np.sort(rng.integers(0, len(pts), size=near))
        pts = np.ascontiguousarray(np.insert(pts, pos, d, axis=0))
        ring = np.ascontiguousarray(np.insert(ring, pos, rng.integers(0, 16, size=near), axis=0))
    ps = syn.tracking_particles(256, gt, sigma_xy=0.05, sigma_z=0.05, sigma_yaw=0.03)
    rc_c, cpu, pose_c, err_c, _ = shim.evaluate_cloud(ev, ps, pts, ring, use_cuda=False)
    rc_g, gpu, pose_g, err_g, used = shim.evaluate_cloud(ev, ps, pts, ring, use_cuda=True, desync=True, n_rings=64)
    assert rc_c == 0 and rc_g == 0, (err_c, err_g)
    want, _ = Oracle().reduce_scan(pts, ring, 0.064, n_rings=64, ring_desync=True)
    assert used == len(want) and 0 < used < len(pts)
    rel = common.rel_err(gpu[:, 6], cpu[:, 6])
    assert rel.max() <= 1e-5, f"normalised weights: max rel err {rel.max():.2e}"
    np.testing.assert_allclose(pose_g[3:], pose_c[3:], atol=1e-5)
    if near:
        # product default (every point keeps its own ring) differs from the reference's desynchronised pairing
        rc_f, fixed, _, err_f, used_f = shim.evaluate_cloud(ev, ps, pts, ring, use_cuda=True, desync=False, n_rings=64)
        assert rc_f == 0, err_f
        want_f, _ = Oracle().reduce_scan(pts, ring, 0.064, n_rings=64, ring_desync=False)
        assert used_f == len(want_f)
    shim.eval_destroy(ev)


@pytest.mark.gpu
def test_reference_facade_over_several_devices(shim, small_map, monkeypatch):
    """TSDFLOC_DEVICES="a,b,..": the reference's TSDFEvaluator::evaluate(use_cuda=true) runs sharded over those GPUs from the one
    process (here: the same device twice, or two devices when present) and returns what the single-device run returns."""
    import torch
    ps, pts = _workload()
    ev = shim.eval_create(small_map)
    rc, one, pose_one, err = shim.evaluate(ev, ps, pts, syn.CALIB_TF, use_cuda=True)
    assert rc == 0, err
    shim.eval_destroy(ev)
    monkeypatch.setenv("TSDFLOC_DEVICES", "0,1" if torch.cuda.device_count() >= 2 else "0,0")
    ev = shim.eval_create(small_map)
    assert ev, shim.last_error()
    rc, many, pose_many, err = shim.evaluate(ev, ps, pts, syn.CALIB_TF, use_cuda=True)
    assert rc == 0, err
    assert many.tobytes() == one.tobytes() and pose_many.tobytes() == pose_one.tobytes()
    m_ref, out_ref, _ = shim.systematic_resample(many, 3)
    m_gpu, out_gpu = shim.gpu_systematic_resample(many, 3)      # the resampler borrows rank 0's context
    assert m_gpu == m_ref and np.array_equal(out_gpu, out_ref)
    shim.eval_destroy(ev)


@pytest.mark.gpu
def test_reference_evaluateParticles_over_several_devices(shim, small_map, monkeypatch):
    """The call mcl_3d actually makes — TSDFEvaluator::evaluateParticles(cloud) — with TSDFLOC_DEVICES naming several GPUs:
    the first device reduces the cloud, every device evaluates its particle slice (tsdfloc_multi_sensor_update_cloud). Same
    bytes as the single-device run, same reduced scan size."""
    import torch
    gt = (0.4, -0.3, 1.2, 0.01, -0.02, 0.4)
    pts, ring = syn.make_scan("vlp16", gt, room_lo=(-3.0, -2.5, 0.0), room_hi=(3.0, 2.5, 3.0))
    ps = syn.tracking_particles(301, gt, sigma_xy=0.05, sigma_z=0.05, sigma_yaw=0.03)
    ev = shim.eval_create_cell(small_map, 0.064)
    assert ev, shim.last_error()
    rc, one, pose_one, err, used_one = shim.evaluate_cloud(ev, ps, pts, ring, use_cuda=True, desync=False, n_rings=64)
    assert rc == 0, err
    shim.eval_destroy(ev)
    monkeypatch.setenv("TSDFLOC_DEVICES", "0,1,0" if torch.cuda.device_count() >= 2 else "0,0,0")
    ev = shim.eval_create_cell(small_map, 0.064)
    assert ev, shim.last_error()
    for _ in range(2):
        rc, many, pose_many, err, used_many = shim.evaluate_cloud(ev, ps, pts, ring, use_cuda=True, desync=False, n_rings=64)
        assert rc == 0, err
        assert used_many == used_one and 0 < used_one < len(pts)
        assert many.tobytes() == one.tobytes() and pose_many.tobytes() == pose_one.tobytes()
    shim.eval_destroy(ev)
