"""Scan reduction: the C oracle pinned against the reference's own evaluateParticles reduction
(src/evaluation/tsdf_evaluator.cpp:304-376, run verbatim through oracle/_ref) and against the numpy generator the
workloads use. CPU only."""
import numpy as np
import pytest

from oracle_lib import Oracle, Ref, ref_available
from tsdf_localization_b200 import synthetic as syn

needs_ref = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")


def scan_with_rings(kind="vlp16", seed=3, near=0, shuffle=False):
    """A ray-cast scan (ring-major) plus `near` points closer than 1 m spliced in at random positions."""
    pts, ring = syn.make_scan(kind, syn.GT_POSE)
    rng = np.random.default_rng(seed)
    if shuffle:   # azimuth-major like a spinning-LiDAR driver: rings interleaved
        perm = rng.permutation(len(pts))
        pts, ring = pts[perm], ring[perm]
    if near:
        d = rng.normal(size=(near, 3)).astype(np.float32)
        d *= (rng.uniform(0.05, 0.99, size=(near, 1)) / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        pos = np.sort(rng.integers(0, len(pts), size=near))
        pts = np.insert(pts, pos, d, axis=0)
        ring = np.insert(ring, pos, rng.integers(0, int(ring.max()) + 1, size=near), axis=0)
    return np.ascontiguousarray(pts, dtype=np.float32), np.ascontiguousarray(ring, dtype=np.int32)


@needs_ref
@pytest.mark.parametrize("cell", [0.064, 0.256, 0.05])
@pytest.mark.parametrize("shuffle", [False, True])
def test_oracle_matches_reference_without_near_points(cell, shuffle):
    pts, ring = scan_with_rings(shuffle=shuffle)
    want = Ref().reduce_scan(pts, ring, cell)
    got, src = Oracle().reduce_scan(pts, ring, cell, n_rings=64)
    assert got.tobytes() == want.tobytes()
    assert np.array_equal(pts[src], got)
    # without near points the reference's ring iterator stays in step: both modes agree
    got2, _ = Oracle().reduce_scan(pts, ring, cell, n_rings=64, ring_desync=True)
    assert got2.tobytes() == want.tobytes()


@needs_ref
@pytest.mark.parametrize("shuffle", [False, True])
def test_oracle_reproduces_the_ring_iterator_desync(shuffle):
    """Dropped (< 1 m) points do not advance the reference's ring iterator (tsdf_evaluator.cpp:319-322)."""
    pts, ring = scan_with_rings(near=500, shuffle=shuffle)
    want = Ref().reduce_scan(pts, ring, 0.064)
    got, _ = Oracle().reduce_scan(pts, ring, 0.064, n_rings=64, ring_desync=True)
    assert got.tobytes() == want.tobytes()
    fixed, src = Oracle().reduce_scan(pts, ring, 0.064, n_rings=64, ring_desync=False)
    assert np.array_equal(pts[src], fixed)
    assert (np.linalg.norm(fixed.astype(np.float64), axis=1) >= 1.0 - 1e-6).all()
    if not shuffle:
        # product mode: every point keeps its own ring -> output is ring-major and per ring in cloud order
        r = ring[src]
        assert (np.diff(r) >= 0).all()
        assert all((np.diff(src[r == k]) > 0).all() for k in np.unique(r))


@needs_ref
def test_oracle_matches_reference_64_rings_negative_coordinates():
    rng = np.random.default_rng(11)
    n = 40000
    pts = (rng.uniform(-30, 30, size=(n, 3)) * np.array([1, 1, 0.2])).astype(np.float32)
    pts[:50] = pts[50:100]                      # exact duplicates
    pts[100:150] = pts[150:200] + np.float32(1e-4)   # near duplicates, mostly the same cell
    ring = rng.integers(0, 64, size=n).astype(np.int32)
    want = Ref().reduce_scan(pts, ring, 0.256)
    got, _ = Oracle().reduce_scan(pts, ring, 0.256, n_rings=64, ring_desync=True)
    assert got.tobytes() == want.tobytes()


def test_oracle_matches_numpy_generator():
    pts, ring = scan_with_rings("os1-128")
    want, want_ring = syn.reduce_scan(pts, ring, 0.256)
    got, src = Oracle().reduce_scan(pts, ring, 0.256, n_rings=128)
    assert got.tobytes() == want.tobytes()
    assert np.array_equal(ring[src], want_ring)


def test_oracle_edge_cases():
    o = Oracle()
    out, src = o.reduce_scan(np.zeros((0, 3), np.float32), np.zeros(0, np.int32), 0.064)
    assert len(out) == 0 and len(src) == 0
    pts = np.array([[2, 0, 0], [np.nan, 0, 0], [0.5, 0, 0], [2.001, 0, 0], [np.inf, 1, 1], [2, 0, 0]], dtype=np.float32)
    ring = np.array([3, 3, 3, 3, 3, 4], dtype=np.int32)
    out, src = o.reduce_scan(pts, ring, 0.064, n_rings=8)
    assert src.tolist() == [0, 5]               # NaN/inf dropped, < 1 m dropped, same (ring, cell) deduplicated, ring 4 kept
    with pytest.raises(ValueError):
        o.reduce_scan(pts, ring, 0.064, n_rings=4)
    with pytest.raises(ValueError):
        o.reduce_scan(pts, np.array([0, 0, 0, -1, 0, 0], np.int32), 0.064, n_rings=4)


def test_centre_variant_properties():
    """cuda_evaluator.cu:96-108 / num_particles_eval.cpp:138-146: cell centres in double arithmetic, duplicates removed."""
    pts, ring = scan_with_rings()
    o = Oracle()
    out, src = o.reduce_scan_centres(pts, None, 0.064)
    want = np.floor(pts.astype(np.float64) / 0.064) * 0.064 + 0.032
    assert np.array_equal(out, want[src].astype(np.float32))
    assert len(np.unique(out, axis=0)) == len(out)
    assert len(np.unique(want.astype(np.float32), axis=0)) == len(out)
    out_r, src_r = o.reduce_scan_centres(pts, ring, 0.064, n_rings=16)
    assert len(out_r) >= len(out) and (np.diff(ring[src_r]) >= 0).all()


@needs_ref
def test_oracle_evaluateParticles_end_to_end_matches_reference():
    """Reference evaluateParticles (reduction + CPU evaluation, verbatim) == oracle reduction followed by the oracle evaluation."""
    import common
    from oracle_lib import NEG_REF_HOST_X86
    ref, o = Ref(), Oracle()
    spec, m = common.box_room(small=True)
    rm = ref.map_create(spec.min, spec.max, spec.resolution, spec.init_value)
    assert ref.map_set_data(rm, spec.cells) == 0
    ev = ref.eval_create_cell(rm, 0.256)
    gt = (0.4, -0.3, 1.2, 0.01, -0.02, 0.4)
    pts, ring = syn.make_scan("vlp16", gt, room_lo=(-3.0, -2.5, 0.0), room_hi=(3.0, 2.5, 3.0))
    ps = syn.tracking_particles(64, gt, sigma_xy=0.05, sigma_z=0.05, sigma_yaw=0.03)
    rc, got, pose, err, _ = ref.evaluate_cloud(ev, ps, pts, ring, use_cuda=False)
    assert rc == 0, err
    red, _ = o.reduce_scan(pts, ring, 0.256, n_rings=64, ring_desync=True)
    want = o.evaluate(common.oracle_map_of(o, m), common.DEFAULT_PARAMS, ps, red, syn.IDENTITY_TF, mode=NEG_REF_HOST_X86)
    assert want["status"] == 0
    assert common.rel_err(got[:, 6], want["particles"][:, 6]).max() <= 2e-6   # weight_sum order differs (OpenMP reduction)
    ref.eval_destroy(ev)
    ref.map_destroy(rm)
