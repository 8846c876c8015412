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


MUTANTS = ["1e3", "+5", "-0", "0x10", "nan", "inf", "-inf", "1.5.2", "abc", "1e999", "-1e999", "1e-999", ".5", "5.", "-", "+", "1e", "1e+",
           "3,5", "7f", "0007", "1E2", "-.25e-1", "2147483648", "-2147483649", "99999999999999999999", "3.9999999999", "", "1 2", "\t8\n",
           "+inf", "-nan", "-0x10", "+0x1p3", "1e+5x", "-infinity", "--1", "+-1", "1..2", "1.e3", ".e3", "e3", "1e3e4", "1e-", "-.", "+.5",
           "-5.", "0e0", "00.10", "1d5"]
SAFE_FOR_COUNTS = ("+5", "abc", "", "-", "0x10", "1.5.2", "0007", "3,5", "1 2", "-0", "+0x1p3", "--1", "+-1", "-.")   # others make the reference resize() wildly


@needs_ref
def test_reader_agrees_with_reference_reader_on_mutated_files(tmp_path):
    """Differential test of tsdfloc_mcl_read against the verbatim MCLFile::read (`istream >> size_t / int / float`,
    src/util/mcl_file.cpp:66-113): a small valid snapshot with one token replaced, dropped, doubled, or the file truncated —
    both readers must accept or reject alike, and where they accept, parse the same values."""
    ref = Ref()
    pts, ring, ps, tf, pose = snapshot(3, 2)
    good = tmp_path / "good.mcl"
    MCLFile(good).write(pts, ring, ps, tf, *pose)
    tokens = good.read_text().split()
    count_slots = {0, 1 + 3 * 3 + 3}                     # the two element counts
    cases = []
    for slot in range(len(tokens)):
        for m in MUTANTS:
            if slot in count_slots and m not in SAFE_FOR_COUNTS:
                continue
            t = list(tokens)
            t[slot] = m
            cases.append(" ".join(t))
    for cut in range(0, len(good.read_text()), 7):
        cases.append(good.read_text()[:cut])
    cases.append(good.read_text() + " 1 2 3")                     # trailing tokens are ignored by both
    cases.append(good.read_text().replace("\n", "\r\n"))
    n_ok = 0
    for k, text in enumerate(cases):
        f = tmp_path / "case.mcl"
        f.write_text(text)
        try:
            theirs = ref.mcl_read(f)
        except RuntimeError:
            theirs = None
        try:
            ours = MCLFile(f).read()
        except ValueError:
            ours = None
        assert (ours is None) == (theirs is None), f"case {k}: product {'rejects' if ours is None else 'accepts'}, reference does not: {text[:120]!r}"
        if ours is not None:
            n_ok += 1
            for a, b in zip((ours.points, ours.rings, ours.particles, ours.tf, ours.pose), theirs):
                assert np.asarray(a).tobytes() == np.asarray(b).tobytes(), f"case {k}: {text[:120]!r}"
    assert 0 < n_ok < len(cases)
