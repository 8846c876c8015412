"""``.mcl`` snapshot files (tsdf_localization_b200/mcl_file.py) against the reference's own MCLFile
(src/util/mcl_file.cpp:14-113, compiled verbatim into oracle/_ref): files written by either side are byte-identical and
read back identically by the other. CPU only."""
import numpy as np
import pytest

from oracle_lib import Ref, ref_available
from tsdf_localization_b200 import synthetic as syn
from tsdf_localization_b200.mcl_file import MCLFile

needs_ref = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")


def snapshot(n_points=700, n_particles=300):
    pts, ring = syn.make_scan("vlp16", syn.GT_POSE, n_points=n_points)
    ps = syn.tracking_particles(n_particles, syn.GT_POSE)
    ps[:, 6] = np.random.default_rng(0).random(n_particles).astype(np.float32) * 1e-3
    ps[0, 6] = 0.0
    ps[1, :3] = (1e-7, -123456.789, 3.0e10)          # exponents, rounding to 6 significant digits
    pose = np.array([1.5, -2.25, 0.125, 1.0, 0.0, 0.0, 0.0], dtype=np.float32)
    return pts, ring.astype(np.int32), ps, syn.CALIB_TF.astype(np.float32), pose


@needs_ref
def test_write_is_byte_identical_and_cross_readable(tmp_path):
    ref = Ref()
    pts, ring, ps, tf, pose = snapshot()
    ours, theirs = tmp_path / "ours.mcl", tmp_path / "theirs.mcl"
    MCLFile(ours).write(pts, ring, ps, tf, *pose)
    ref.mcl_write(theirs, pts, ring, ps, tf, pose)
    assert ours.read_bytes() == theirs.read_bytes()
    got = MCLFile(theirs).read()
    r_pts, r_ring, r_ps, r_tf, r_pose = ref.mcl_read(ours)
    assert got.points.tobytes() == r_pts.tobytes() and np.array_equal(got.rings, r_ring)
    assert got.particles.tobytes() == r_ps.tobytes() and got.tf.tobytes() == r_tf.tobytes() and got.pose.tobytes() == r_pose.tobytes()
    # the text keeps 6 significant digits (ostream default): values survive to that precision
    np.testing.assert_allclose(got.points, pts, rtol=1e-5, atol=0)
    np.testing.assert_allclose(got.particles, ps, rtol=1e-5, atol=0)


def test_round_trip_and_errors(tmp_path):
    pts, ring, ps, tf, pose = snapshot(50, 20)
    f = tmp_path / "s.mcl"
    MCLFile(f).write(pts, ring, ps, tf, *pose)
    a = MCLFile(f).read()
    MCLFile(f).write(a.points, a.rings, a.particles, a.tf, *a.pose)
    b = MCLFile(f).read()
    assert a.points.tobytes() == b.points.tobytes() and a.particles.tobytes() == b.particles.tobytes()   # %g is idempotent
    empty = tmp_path / "e.mcl"
    MCLFile(empty).write(np.zeros((0, 3)), np.zeros(0, int), np.zeros((0, 7)), tf, *pose)
    e = MCLFile(empty).read()
    assert e.points.shape == (0, 3) and e.particles.shape == (0, 7)
    (tmp_path / "bad.mcl").write_text("3\n1 2 3\n")
    with pytest.raises(ValueError, match="Could not read mcl data"):
        MCLFile(tmp_path / "bad.mcl").read()
    with pytest.raises(OSError):
        MCLFile(tmp_path / "missing.mcl").read()
