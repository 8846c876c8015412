"""GPU scan reduction (tsdfloc_reduce_scan*, tsdfloc_sensor_update_cloud) against the CPU oracle, which tests/test_reduce_oracle.py
pins against the reference's own evaluateParticles reduction (src/evaluation/tsdf_evaluator.cpp:304-376). Bar: bit-exact
points in identical order."""
import ctypes as C

import numpy as np
import pytest

import common
from oracle_lib import NEG_AS_MISS
from test_reduce_oracle import scan_with_rings
from tsdf_localization_b200 import CudaEvaluator, TSDFEvaluator, capi, synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def evaluator():
    _, m = common.box_room(small=True)
    ev = CudaEvaluator(m)
    yield ev
    ev.close()


@pytest.mark.parametrize("kind,n_rings", [("vlp16", 16), ("os1-128", 128)])
@pytest.mark.parametrize("cell", [0.064, 0.256, 0.05])
@pytest.mark.parametrize("shuffle", [False, True])
def test_reduce_matches_oracle(oracle, evaluator, kind, n_rings, cell, shuffle):
    pts, ring = scan_with_rings(kind, near=300, shuffle=shuffle)
    for desync in (False, True):
        want, want_src = oracle.reduce_scan(pts, ring, cell, n_rings=n_rings, ring_desync=desync)
        got, src = evaluator.reduce_scan(pts, ring, cell, n_rings=n_rings, ring_desync_like_reference=desync, want_src=True)
        assert got.shape == want.shape, f"{len(got)} points, oracle {len(want)}"
        assert np.array_equal(src, want_src)
        assert got.tobytes() == want.tobytes()
    assert 0 < len(got) < len(pts)


def test_reduce_int16_rings_and_strided_cloud(oracle, evaluator, lib):
    """A PointCloud2-style byte buffer: x y z at offset 0, ring (int16, as the reference reads it) at offset 16, point_step 24."""
    pts, ring = scan_with_rings("vlp16", near=100, shuffle=True)
    n = len(pts)
    cloud = np.zeros((n, 24), dtype=np.uint8)
    cloud[:, 0:12] = pts.view(np.uint8).reshape(n, 12)
    cloud[:, 16:18] = ring.astype(np.int16).view(np.uint8).reshape(n, 2)
    out = np.empty((n, 3), dtype=np.float32)
    src = np.empty(n, dtype=np.uint32)
    n_out = C.c_uint64(0)
    base = cloud.ctypes.data
    rc = lib.tsdfloc_reduce_scan(evaluator.ctx, C.c_void_p(base), 24, C.c_void_p(base + 16), 24, 2, n, C.c_double(0.064), 16, 0,
                                 out.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), n, C.byref(n_out))
    capi.check(lib, evaluator.ctx, rc)
    want, want_src = oracle.reduce_scan(pts, ring, 0.064, n_rings=16)
    m = int(n_out.value)
    assert m == len(want) and np.array_equal(src[:m], want_src) and out[:m].tobytes() == want.tobytes()


def test_reduce_centres_variant(oracle, evaluator):
    pts, ring = scan_with_rings("vlp16")
    for rg in (None, ring):
        want, want_src = oracle.reduce_scan_centres(pts, rg, 0.064, n_rings=16)
        got, src = evaluator.reduce_scan(pts, rg, 0.064, n_rings=16, emit_centres=True, want_src=True)
        assert np.array_equal(src, want_src) and got.tobytes() == want.tobytes()


def test_reduce_edge_cases(oracle, evaluator, lib):
    # empty cloud
    out = evaluator.reduce_scan(np.zeros((0, 3), np.float32), np.zeros(0, np.int32))
    assert out.shape == (0, 3)
    # everything dropped (all nearer than 1 m / non-finite)
    pts = np.array([[0.1, 0.2, 0.3], [np.nan, 0, 0], [np.inf, 0, 0], [0.5, 0.5, 0.5]], dtype=np.float32)
    assert len(evaluator.reduce_scan(pts, np.zeros(4, np.int32))) == 0
    # one point, not a multiple of anything
    one = np.array([[3.0, 1.0, 0.5]], dtype=np.float32)
    got = evaluator.reduce_scan(one, np.array([7], np.int32), n_rings=8)
    assert got.tobytes() == one.tobytes()
    # ring out of range -> E_BAD_ARG (the reference is out of bounds there)
    with pytest.raises(capi.TsdflocError) as e:
        evaluator.reduce_scan(one, np.array([8], np.int32), n_rings=8)
    assert e.value.status == capi.E_BAD_ARG
    with pytest.raises(capi.TsdflocError):
        evaluator.reduce_scan(one, np.array([-1], np.int32), n_rings=8)
    with pytest.raises(capi.TsdflocError):
        evaluator.reduce_scan(one, np.array([0], np.int32), cell_size=0.0)
    with pytest.raises(capi.TsdflocError):
        evaluator.reduce_scan(one, np.array([0], np.int32), n_rings=4096)
    # all points in one cell of one ring -> the first survives; many duplicates stress the atomicMin path
    dup = np.tile(np.array([[5.01, 5.01, 1.01]], dtype=np.float32), (70000, 1))
    dup[1:] += np.float32(1e-3)
    got, src = evaluator.reduce_scan(dup, np.zeros(len(dup), np.int32), 0.256, want_src=True)
    assert src.tolist() == [0] and got.tobytes() == dup[:1].tobytes()
    # 1,024 rings (the maximum), random cloud, collisions across rings
    rng = np.random.default_rng(5)
    pts = rng.uniform(-8, 8, size=(50000, 3)).astype(np.float32)
    ring = rng.integers(0, 1024, size=len(pts)).astype(np.int32)
    want, want_src = oracle.reduce_scan(pts, ring, 0.5, n_rings=1024)
    got, src = evaluator.reduce_scan(pts, ring, 0.5, n_rings=1024, want_src=True)
    assert np.array_equal(src, want_src) and got.tobytes() == want.tobytes()


def test_reduce_idempotent_and_full_size(oracle, evaluator):
    """Size-independent properties at the OS1-128 size: reducing a reduced scan changes nothing; output is ring-major, in
    cloud order inside a ring, one point per (ring, cell)."""
    pts, ring = scan_with_rings("os1-128", near=1000)
    red, src = evaluator.reduce_scan(pts, ring, 0.256, want_src=True)
    r = ring[src]
    assert (np.diff(r) >= 0).all()
    assert all((np.diff(src[r == k]) > 0).all() for k in np.unique(r))
    cells = np.floor(red / np.float32(0.256)).astype(np.int64)
    keys = np.concatenate([r[:, None].astype(np.int64), cells], axis=1)
    assert len(np.unique(keys, axis=0)) == len(keys)
    again, src2 = evaluator.reduce_scan(red, r, 0.256, want_src=True)
    assert again.tobytes() == red.tobytes() and np.array_equal(src2, np.arange(len(red)))


def test_sensor_update_cloud_equals_reduce_then_update(oracle):
    """tsdfloc_sensor_update_cloud == oracle reduction followed by the oracle evaluation (evaluateParticles end to end)."""
    spec, m = common.box_room(small=True)
    om = common.oracle_map_of(oracle, m)
    gt = (0.4, -0.3, 1.2, 0.01, -0.02, 0.4)
    pts, ring = syn.make_scan("vlp16", gt, room_lo=(-3.0, -2.5, 0.0), room_hi=(3.0, 2.5, 3.0))
    ps = syn.tracking_particles(300, gt, sigma_xy=0.2)
    red, _ = oracle.reduce_scan(pts, ring, 0.064, n_rings=16)
    ref = oracle.evaluate(om, common.DEFAULT_PARAMS, ps, red, syn.CALIB_TF, mode=NEG_AS_MISS)
    te = TSDFEvaluator(m, reduction_cell_size=0.064)
    mine = ps.copy()
    pose = te.evaluateParticles(mine, pts, ring, syn.CALIB_TF, n_rings=16)
    assert common.rel_err(mine[:, 6], ref["particles"][:, 6]).max() <= 1e-5
    assert np.allclose(pose.position, ref["mean"][:3], atol=1e-4)
    # the resident particle set can be resampled right away
    out = te.cuda_evaluator_.resample_systematic(0.37 / len(ps), capacity=len(ps) + 64)
    m_ref, parents_ref = oracle.systematic_resample(mine[:, 6], 0.37 / len(ps))
    assert len(out) == m_ref and np.array_equal(out[:, :6], mine[parents_ref][:, :6])
    # identical to evaluate() on the pre-reduced scan
    mine2 = ps.copy()
    te.evaluate(mine2, red, syn.CALIB_TF)
    assert mine2.tobytes() == mine.tobytes()
    te.cuda_evaluator_.close()
