"""The dense voxel layout and the certified speculative voxel index of the evaluation kernel (tsdfloc_device.cuh,
tsdfloc_eval.cuh): every combination of layout and index path must produce the reference's flat voxel indices, hit counts
and sequential fp32 sums BIT FOR BIT — including the inputs built to sit exactly on voxel faces, the non-finite scans and the
particles whose certified margin is too wide, all of which must fall back to the exact path on their own.
Checker: the CPU oracle (pinned against the unmodified reference)."""
import numpy as np
import pytest

import common
from oracle_lib import NEG_AS_MISS
from tsdf_localization_b200 import CudaEvaluator, CudaSubVoxelMap, capi, synthetic as syn

pytestmark = pytest.mark.gpu

MODES = [("dense+speculative", -1, -1), ("dense, exact index", -1, 0), ("brick layout", 0, 0)]


@pytest.fixture(scope="module")
def room():
    return common.box_room()


@pytest.fixture(scope="module")
def ev(room):
    e = CudaEvaluator(room[1])
    yield e
    e.close()


@pytest.fixture(scope="module")
def omap(oracle, room):
    return common.oracle_map_of(oracle, room[1])


def _all_modes(ev, ps, pts, tf, want_idx=True):
    out = {}
    try:
        for name, dense, spec in MODES:
            ev.tune(capi.TUNE_DENSE, dense)
            ev.tune(capi.TUNE_SPECULATE, spec)
            for pairing in (1, 2):
                ev.tune(capi.TUNE_EVAL_PAIRING, pairing)
                out[(name, pairing)] = ev.debug_eval(ps, pts, tf, want_idx=want_idx)
    finally:
        ev.tune(capi.TUNE_DENSE, -1)
        ev.tune(capi.TUNE_SPECULATE, -1)
        ev.tune(capi.TUNE_EVAL_PAIRING, 0)
    return out


def _check(ref, got):
    for key, (idx, hits, raw) in got.items():
        if idx is not None:
            bad = int((idx != ref["idx"]).sum())
            assert bad == 0, f"{key}: {bad} of {idx.size} flat voxel indices differ"
        assert np.array_equal(hits, ref["hits"]), f"{key}: hit counts differ"
        assert raw.tobytes() == ref["raw"].tobytes(), f"{key}: raw weights differ"


def test_proofs_passed_and_path_used(ev):
    st0 = ev.spec_stats()
    assert st0["proven"], "box room at 5 cm: dense layout + speculative index must be available"
    ps, pts, _ = common.config_c2(512)
    mine = ps.copy()
    ev.evaluate(mine, pts, syn.CALIB_TF)
    st = ev.spec_stats()
    steps, redone = st["steps"] - st0["steps"], st["redone"] - st0["redone"]
    assert steps > 0 and st["warps_not_eligible"] == st0["warps_not_eligible"]
    # a step is 64 evaluations x 3 axes; each is uncertain with probability ~ 2 delta / res ~ 5e-4
    assert 0 < redone < 0.3 * steps, f"{redone} of {steps} steps redone"


@pytest.mark.parametrize("tf", [syn.IDENTITY_TF, syn.CALIB_TF], ids=["identity_tf", "calib_tf"])
def test_c1_every_mode_bit_exact(oracle, omap, ev, tf):
    ps, pts, _ = common.config_c1()
    ref = oracle.evaluate(omap, common.DEFAULT_PARAMS, ps, pts, tf, mode=NEG_AS_MISS, want_idx=True)
    _check(ref, _all_modes(ev, ps, pts, tf))


def test_c2_slice_every_mode_bit_exact(oracle, omap, ev):
    ps, pts, _ = common.config_c2(768)
    ref = oracle.evaluate(omap, common.DEFAULT_PARAMS, ps, pts, syn.CALIB_TF, mode=NEG_AS_MISS, want_idx=True)
    _check(ref, _all_modes(ev, ps, pts, syn.CALIB_TF))


def test_points_on_voxel_faces(oracle, omap, ev, room):
    """Axis-aligned particles at lattice positions against points whose coordinates are exact multiples of the voxel size:
    almost every lookup sits ON a voxel face, where the speculative index must give up and the exact path decide."""
    rng = np.random.default_rng(5)
    n, p = 64, 2048
    ps = np.zeros((n, 7), np.float32)
    ps[:, 0] = rng.integers(-60, 60, n) * 0.05
    ps[:, 1] = rng.integers(-60, 60, n) * 0.05
    ps[:, 2] = rng.integers(10, 60, n) * 0.05
    ps[::3, 5] = np.float32(np.pi / 2)      # some quarter turns: cos is 6e-8 off zero, still on-lattice to 1e-6
    pts = (rng.integers(-160, 160, (p, 3)) * 0.05).astype(np.float32)
    pts[:, 2] = (rng.integers(-20, 60, p) * 0.05).astype(np.float32)
    ref = oracle.evaluate(omap, common.DEFAULT_PARAMS, ps, pts, syn.IDENTITY_TF, mode=NEG_AS_MISS, want_idx=True)
    st0 = ev.spec_stats()
    _check(ref, _all_modes(ev, ps, pts, syn.IDENTITY_TF))
    st = ev.spec_stats()
    assert st["redone"] - st0["redone"] > 0.5 * (st["steps"] - st0["steps"]), "these inputs should defeat the speculation"


def test_points_one_ulp_around_faces(oracle, omap, ev):
    """Offsets one and two ulps to either side of voxel and cell faces, for a translation-only particle: the exact rounding of
    the reference decides which voxel it is."""
    base = np.arange(0, 400, dtype=np.float64) * 0.05 - 10.0
    xs = []
    for k in (-2, -1, 0, 1, 2):
        v = base.astype(np.float32)
        for _ in range(abs(k)):
            v = np.nextafter(v, np.float32(np.inf if k > 0 else -np.inf))
        xs.append(v)
    x = np.concatenate(xs)
    pts = np.stack([x, np.roll(x, 7), np.abs(np.roll(x, 13)) * 0.25], axis=1).astype(np.float32)
    ps = np.zeros((6, 7), np.float32)
    ps[:, 0] = [0.0, 0.05, 1.0, -0.35, 2.5e-6, 0.025]
    ps[:, 1] = [0.0, -0.05, 0.5, 0.15, -2.5e-6, 0.0]
    ps[:, 2] = [0.0, 0.0, 0.05, 0.1, 0.0, 1.0]
    ref = oracle.evaluate(omap, common.DEFAULT_PARAMS, ps, pts, syn.IDENTITY_TF, mode=NEG_AS_MISS, want_idx=True)
    _check(ref, _all_modes(ev, ps, pts, syn.IDENTITY_TF))


def test_outside_the_map_and_negative_band(oracle, omap, ev):
    """Particles near and beyond every face of the bounding box: clamped coordinates, the band below map.min, the last voxel."""
    rng = np.random.default_rng(11)
    pts, _ = syn.make_scan("vlp16", syn.GT_POSE, n_points=3000)
    ps = syn.tracking_particles(96, syn.GT_POSE, sigma_xy=6.0, sigma_z=2.5, sigma_yaw=3.0, seed=3)
    ps[:8, 0] += 30.0
    ps[8:16, 1] -= 30.0
    ps[16:24, 2] += 8.0
    ps[24:32, 2] -= 8.0
    ref = oracle.evaluate(omap, common.DEFAULT_PARAMS, ps, pts, syn.CALIB_TF, mode=NEG_AS_MISS, want_idx=True)
    _check(ref, _all_modes(ev, ps, pts, syn.CALIB_TF))
    assert (ref["idx"] == ref["idx"].max()).mean() > 0.2     # plenty of misses in this workload


def test_far_particles_are_not_speculated(oracle, omap, ev):
    """A particle 10^6 m away has a margin wider than a voxel: its warp must run the exact loop (and still match), next to
    eligible warps in the same launch."""
    pts, _ = syn.make_scan("vlp16", syn.GT_POSE, n_points=1500)
    ps = syn.tracking_particles(40, syn.GT_POSE, seed=9)
    ps[4, 0] = 1.0e6
    ps[11, 1] = -3.0e5
    ps[30, :3] = (2.0e4, 2.0e4, 50.0)
    ref = oracle.evaluate(omap, common.DEFAULT_PARAMS, ps, pts, syn.IDENTITY_TF, mode=NEG_AS_MISS, want_idx=True)
    st0 = ev.spec_stats()
    _check(ref, _all_modes(ev, ps, pts, syn.IDENTITY_TF))
    st = ev.spec_stats()
    assert st["warps_not_eligible"] > st0["warps_not_eligible"] and st["steps"] > st0["steps"]


def test_non_finite_scan_disables_speculation(oracle, omap, ev):
    pts, _ = syn.make_scan("vlp16", syn.GT_POSE, n_points=1200)
    pts = pts.copy()
    pts[17] = (np.nan, 1.0, 0.5)
    pts[400] = (np.inf, -np.inf, 0.0)
    pts[401] = (3.0e38, 3.0e38, 3.0e38)
    ps = syn.tracking_particles(24, syn.GT_POSE, seed=2)
    ref = oracle.evaluate(omap, common.DEFAULT_PARAMS, ps, pts, syn.CALIB_TF, mode=NEG_AS_MISS, want_idx=True)
    st0 = ev.spec_stats()
    _check(ref, _all_modes(ev, ps, pts, syn.CALIB_TF))
    st = ev.spec_stats()
    assert st["steps"] == st0["steps"], "a scan with non-finite points must not be speculated"
    # and the next, finite scan is speculated again (the bound is per scan)
    pts2, _ = syn.make_scan("vlp16", syn.GT_POSE, n_points=1200)
    mine = ps.copy()
    ev.evaluate(mine, pts2, syn.CALIB_TF)
    assert ev.spec_stats()["steps"] > st["steps"]


def test_budget_and_unaligned_resolution(oracle, room):
    """dense_budget_bytes = 1 keeps the brick layout; a 6.4 cm map (sub_dim * res != 1: the reference's own map resolution)
    never gets the dense layout. Both still match the oracle."""
    ps, pts, _ = common.config_c1()
    e = CudaEvaluator(room[1], dense_budget_bytes=1)
    assert not e.spec_stats()["proven"]
    om = common.oracle_map_of(oracle, room[1])
    ref = oracle.evaluate(om, common.DEFAULT_PARAMS, ps, pts, syn.CALIB_TF, mode=NEG_AS_MISS, want_idx=True)
    idx, hits, raw = e.debug_eval(ps, pts, syn.CALIB_TF)
    assert np.array_equal(idx, ref["idx"]) and np.array_equal(hits, ref["hits"]) and raw.tobytes() == ref["raw"].tobytes()
    e.close()
    _, m64 = common.box_room(resolution=0.064)
    e = CudaEvaluator(m64)
    assert not e.spec_stats()["proven"]
    om = common.oracle_map_of(oracle, m64)
    ref = oracle.evaluate(om, common.DEFAULT_PARAMS, ps, pts, syn.CALIB_TF, mode=NEG_AS_MISS, want_idx=True)
    idx, hits, raw = e.debug_eval(ps, pts, syn.CALIB_TF)
    assert np.array_equal(idx, ref["idx"]) and np.array_equal(hits, ref["hits"]) and raw.tobytes() == ref["raw"].tobytes()
    e.close()


@pytest.mark.parametrize("res", [0.1, 0.125, 0.25])
def test_other_aligned_resolutions(oracle, res):
    _, m = common.box_room(resolution=res)
    e = CudaEvaluator(m)
    assert e.spec_stats()["proven"], f"{res} m is an aligned resolution"
    om = common.oracle_map_of(oracle, m)
    ps, pts, _ = common.config_c1()
    ref = oracle.evaluate(om, common.DEFAULT_PARAMS, ps, pts, syn.CALIB_TF, mode=NEG_AS_MISS, want_idx=True)
    _check(ref, _all_modes(e, ps, pts, syn.CALIB_TF))
    e.close()
