"""GPU edge cases through the C ABI: ragged / tiny / empty inputs, non-finite points, the negative-offset policy, other map
resolutions, sensor-model variants, error statuses, several contexts per process. Checker: the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import common
from oracle_lib import NEG_AS_MISS, NEG_REF_DEVICE_SAT, NEG_REF_HOST_X86
from tsdf_localization_b200 import CudaEvaluator, CudaSubVoxelMap, SystematicResampler, capi, likelihood_init, likelihood_value
from tsdf_localization_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

GT = (0.4, -0.3, 1.2, 0.01, -0.02, 0.4)
ROOM = dict(room_lo=(-3.0, -2.5, 0.0), room_hi=(3.0, 2.5, 3.0))


@pytest.fixture(scope="module")
def small():
    spec, m = common.box_room(small=True)
    return spec, m


@pytest.fixture(scope="module")
def ev(small):
    e = CudaEvaluator(small[1])
    yield e
    e.close()


@pytest.fixture(scope="module")
def omap(oracle, small):
    return common.oracle_map_of(oracle, small[1])


def _scan(p):
    pts, _ = syn.make_scan("vlp16", GT, n_points=p, **ROOM)
    return pts


@pytest.mark.parametrize("n,p", [(1, 1), (1, 31), (2, 32), (3, 33), (1, 63), (2, 64), (3, 65), (7, 255), (5, 256), (9, 257), (2, 511),
                                 (2, 512), (5, 513), (33, 1000), (4, 1025), (64, 4097), (257, 511)])
def test_ragged_shapes_bit_exact(oracle, omap, ev, n, p):
    """P not a multiple of the 32- / 64-point step or of the 256- / 512-point summation block; odd particle counts (the
    particle-pair shape pairs particles); both pairings and both register budgets of the evaluation kernel."""
    ps = syn.tracking_particles(n, GT, sigma_xy=0.2, seed=n + p)
    pts = _scan(p)
    ref = oracle.evaluate(omap, common.DEFAULT_PARAMS, ps, pts, syn.CALIB_TF, want_idx=True)
    try:
        for pairing in (1, 2):
            for registers in (1, 2):          # the 64- and the 128-register build of the kernel
                ev.tune(capi.TUNE_EVAL_PAIRING, pairing)
                ev.tune(capi.TUNE_EVAL_REGISTERS, registers)
                idx, hits, raw = ev.debug_eval(ps, pts, syn.CALIB_TF)
                assert np.array_equal(idx, ref["idx"]) and np.array_equal(hits, ref["hits"])
                assert raw.tobytes() == ref["raw"].tobytes()
                mine = ps.copy()
                ev.evaluate(mine, pts, syn.CALIB_TF)
                assert common.rel_err(mine[:, 6], ref["particles"][:, 6]).max() <= 1e-5
    finally:
        ev.tune(capi.TUNE_EVAL_PAIRING, 0)
        ev.tune(capi.TUNE_EVAL_REGISTERS, 0)


@pytest.mark.parametrize("pairing", [1, 2], ids=["particle_pairs", "point_pairs"])
@pytest.mark.parametrize("division", [capi.DIV_IEEE, capi.DIV_THREE, capi.DIV_BRACKET], ids=["ieee", "three", "bracket"])
def test_every_pairing_and_quotient_mode_agrees(oracle, omap, small, pairing, division):
    """The evaluation kernel's two pairings (two particles per warp / two points per lane) and three quotient modes are
    interchangeable: flat indices (dumped by the production kernel itself), hit counts and raw weights equal the oracle's."""
    e = CudaEvaluator(small[1])
    mode, open_brackets = e.division_mode()
    assert mode == capi.DIV_BRACKET and 0 < open_brackets < 2 ** 30 // 1000, "5 cm voxels: the bracket must be proven, and tight"
    e.tune(capi.TUNE_EVAL_PAIRING, pairing)
    e.tune(capi.TUNE_DIVISION, division)
    for n, p in ((301, 3001), (7, 513), (64, 64), (2, 4096)):
        ps = syn.tracking_particles(n, GT, sigma_xy=0.2, seed=n)
        pts = _scan(p)
        ref = oracle.evaluate(omap, common.DEFAULT_PARAMS, ps, pts, syn.CALIB_TF, want_idx=True)
        idx, hits, raw = e.debug_eval(ps, pts, syn.CALIB_TF)
        assert np.array_equal(idx, ref["idx"]) and np.array_equal(hits, ref["hits"])
        assert raw.tobytes() == ref["raw"].tobytes()
        mine = ps.copy()
        e.evaluate(mine, pts, syn.CALIB_TF)      # the non-dumping instantiation
        assert common.rel_err(mine[:, 6], ref["particles"][:, 6]).max() <= 1e-5
    e.close()


def test_open_brackets_are_redone_exactly(oracle):
    """Points placed ON voxel faces (coordinates k * resolution as the map computes them) put sub-voxel quotients within an
    ulp of an integer: the bracketed quotient is open there, the kernel must take the exact division for those blocks
    (statistics say it did) and still match the oracle bit for bit."""
    spec, m = common.box_room(small=True)
    om = common.oracle_map_of(oracle, m)
    e = CudaEvaluator(m)
    rng = np.random.default_rng(5)
    res = np.float32(spec.resolution)
    k = rng.integers(-40, 40, size=(4096, 3)).astype(np.float32)
    pts = (k * res).astype(np.float32)
    pts[:, 2] = np.abs(pts[:, 2])
    pts += np.float32(1.0)                       # keep |p| >= 1 m like a real scan
    ps = np.zeros((64, 7), dtype=np.float32)     # grid-aligned poses: whole-voxel translations, no rotation
    ps[:, :3] = (rng.integers(-10, 10, size=(64, 3)).astype(np.float32) * res)
    ps[:, 2] = np.abs(ps[:, 2])
    for pairing in (1, 2):
        e.tune(capi.TUNE_EVAL_PAIRING, pairing)
        before = e.eval_stats()["redone"]
        ref = oracle.evaluate(om, common.DEFAULT_PARAMS, ps, pts, syn.IDENTITY_TF, want_idx=True)
        idx, hits, raw = e.debug_eval(ps, pts, syn.IDENTITY_TF)
        assert np.array_equal(idx, ref["idx"]) and np.array_equal(hits, ref["hits"]) and raw.tobytes() == ref["raw"].tobytes()
        assert e.eval_stats()["redone"] > before, "no block took the exact path: the test does not exercise open brackets"
    e.close()


def test_non_finite_and_far_points_are_misses(oracle, omap, ev):
    ps = syn.tracking_particles(16, GT, sigma_xy=0.2)
    pts = _scan(200).copy()
    pts[3] = (np.nan, 0.0, 0.0)
    pts[10] = (np.inf, 1.0, 1.0)
    pts[20] = (-np.inf, 1.0, 1.0)
    pts[30] = (1e30, -1e30, 1e30)
    pts[40] = (4e6, 0.0, 0.0)          # beyond the 2^22 range of the magic-number floor: must be clamped, not wrapped
    pts[50] = (-4e6, 5e6, -8e6)
    ref = oracle.evaluate(omap, common.DEFAULT_PARAMS, ps, pts, syn.IDENTITY_TF, want_idx=True)
    idx, hits, raw = ev.debug_eval(ps, pts, syn.IDENTITY_TF)
    assert np.array_equal(idx, ref["idx"])
    data_size = ev.data_size
    for j in (3, 10, 20, 30, 40, 50):
        assert (idx[:, j] == data_size).all()
    assert np.array_equal(hits, ref["hits"])
    # NaN points poison the reference's sum (NaN range test -> a_max; value stays finite here), weights stay comparable
    assert raw.tobytes() == ref["raw"].tobytes()


def test_negative_offset_policy_is_miss(oracle):
    """Map whose walls lie ON the bounding-box faces (margin 0): a quarter of all lookups land at negative axis offsets,
    where the reference is undefined behaviour (its x86 CPU build and its CUDA build disagree). Product == NEG_AS_MISS."""
    spec = syn.box_room_map(likelihood_value, likelihood_init(0.1), resolution=0.05, margin=0.0, **ROOM)
    m = CudaSubVoxelMap(*spec.min, *spec.max, spec.resolution, spec.init_value)
    m.setData(spec.cells)
    om = common.oracle_map_of(oracle, m)
    e = CudaEvaluator(m)
    ps = syn.tracking_particles(64, GT, sigma_xy=0.3)
    pts = _scan(2000)
    pol = oracle.evaluate(om, common.DEFAULT_PARAMS, ps, pts, syn.IDENTITY_TF, mode=NEG_AS_MISS, want_idx=True)
    x86 = oracle.evaluate(om, common.DEFAULT_PARAMS, ps, pts, syn.IDENTITY_TF, mode=NEG_REF_HOST_X86, want_idx=True)
    sat = oracle.evaluate(om, common.DEFAULT_PARAMS, ps, pts, syn.IDENTITY_TF, mode=NEG_REF_DEVICE_SAT, want_idx=True)
    idx, hits, raw = e.debug_eval(ps, pts, syn.IDENTITY_TF)
    assert np.array_equal(idx, pol["idx"]) and raw.tobytes() == pol["raw"].tobytes()
    band_x86, band_sat = (pol["idx"] != x86["idx"]), (pol["idx"] != sat["idx"])
    print(f"negative band: {band_x86.mean():.1%} of pairs differ from the reference's x86 build, {band_sat.mean():.1%} from its CUDA build")
    assert band_x86.mean() > 0.05 and band_sat.mean() > 0.05      # the band is populated: the policy matters here
    # outside the band all three agree with the product
    ok = ~(band_x86 | band_sat)
    assert np.array_equal(idx[ok], x86["idx"][ok]) and np.array_equal(idx[ok], sat["idx"][ok])
    e.close()


def test_negative_offset_saturate_like_reference_gpu(oracle):
    """neg_policy = TSDFLOC_NEG_SATURATE_LIKE_REF_GPU reproduces the reference CUDA evaluator's saturating conversions
    (cuda_eval_particles.h:12-67) on the same margin-0 map, where a quarter of the lookups are in the negative band: flat
    indices, hit counts and raw weights bit-exact against the oracle's restatement of the device semantics — with every
    pairing, incl. NaN / infinite / far points."""
    spec = syn.box_room_map(likelihood_value, likelihood_init(0.1), resolution=0.05, margin=0.0, **ROOM)
    m = CudaSubVoxelMap(*spec.min, *spec.max, spec.resolution, spec.init_value)
    m.setData(spec.cells)
    om = common.oracle_map_of(oracle, m)
    e = CudaEvaluator(m, neg_policy=capi.NEG_SATURATE_LIKE_REF_GPU)
    ps = syn.tracking_particles(65, GT, sigma_xy=0.3)
    pts = _scan(2000).copy()
    pts[3] = (np.nan, 0.0, 0.0)
    pts[10] = (-np.inf, 1.0, 1.0)
    pts[20] = (-4e6, 5e6, -8e6)
    pts[30] = (np.inf, -1.0, 1.0)
    sat = oracle.evaluate(om, common.DEFAULT_PARAMS, ps, pts, syn.IDENTITY_TF, mode=NEG_REF_DEVICE_SAT, want_idx=True)
    miss = oracle.evaluate(om, common.DEFAULT_PARAMS, ps, pts, syn.IDENTITY_TF, mode=NEG_AS_MISS, want_idx=True)
    assert (sat["idx"] != miss["idx"]).mean() > 0.05
    for pairing in (1, 2):
        e.tune(capi.TUNE_EVAL_PAIRING, pairing)
        idx, hits, raw = e.debug_eval(ps, pts, syn.IDENTITY_TF)
        assert np.array_equal(idx, sat["idx"]) and np.array_equal(hits, sat["hits"])
        assert raw.tobytes() == sat["raw"].tobytes()
    e.close()


@pytest.mark.parametrize("res", [0.064, 0.1, 0.03])
def test_other_resolutions(oracle, res):
    """sub_dim 16 / 10 / 34; the exhaustive division self-check decides between the 3-instruction and the IEEE division."""
    spec = syn.box_room_map(likelihood_value, likelihood_init(0.1), resolution=res, **ROOM)
    m = CudaSubVoxelMap(*spec.min, *spec.max, spec.resolution, spec.init_value)
    m.setData(spec.cells)
    om = common.oracle_map_of(oracle, m)
    e = CudaEvaluator(m)
    ps = syn.tracking_particles(100, GT, sigma_xy=0.1, sigma_yaw=0.1)
    pts = _scan(3000)
    ref = oracle.evaluate(om, common.DEFAULT_PARAMS, ps, pts, syn.CALIB_TF, want_idx=True)
    idx, hits, raw = e.debug_eval(ps, pts, syn.CALIB_TF)
    assert np.array_equal(idx, ref["idx"]) and np.array_equal(hits, ref["hits"]) and raw.tobytes() == ref["raw"].tobytes()
    assert hits.sum() > 0.3 * idx.size
    e.close()


@pytest.mark.parametrize("params", [(0.7, 0.2, 0.05, 4.0), (1.0, 0.0, 0.0, 100.0), (0.5, 0.5, 0.3, 3.5)])
def test_sensor_model_variants(oracle, omap, small, params):
    """max_range inside the scan's extent: both branches of the range term (a_range/max_range vs a_max) are exercised."""
    e = CudaEvaluator(small[1], False, *params)
    ps = syn.tracking_particles(50, GT, sigma_xy=0.1)
    pts = _scan(2500)
    r2 = (pts.astype(np.float64) ** 2).sum(1)
    if params[3] < 50:
        assert (r2 < params[3] ** 2).any() and (r2 >= params[3] ** 2).any()
    ref = oracle.evaluate(omap, params, ps, pts, syn.IDENTITY_TF)
    _, hits, raw = e.debug_eval(ps, pts, syn.IDENTITY_TF, want_idx=False)
    assert raw.tobytes() == ref["raw"].tobytes()
    mine = ps.copy()
    e.evaluate(mine, pts, syn.IDENTITY_TF)
    assert common.rel_err(mine[:, 6], ref["particles"][:, 6]).max() <= 1e-5
    e.close()


def test_negative_sensor_model_forces_sequential_sum(oracle, omap, small):
    """a_max < 0 makes addends negative: the integer-block summation is not valid and every block folds sequentially."""
    params = (0.9, 0.1, -0.5, 3.0)
    e = CudaEvaluator(small[1], False, *params)
    ps = syn.tracking_particles(20, GT, sigma_xy=0.1)
    pts = _scan(1500)
    ref = oracle.evaluate(omap, params, ps, pts, syn.IDENTITY_TF)
    _, _, raw = e.debug_eval(ps, pts, syn.IDENTITY_TF, want_idx=False)
    assert raw.tobytes() == ref["raw"].tobytes()
    e.close()


def test_errors_and_states(ev, small):
    ps = syn.tracking_particles(8, GT)
    pts = _scan(64)
    # empty scan: default pose, weights untouched (cuda_evaluator.cu:122-125)
    ps[:, 6] = 0.5
    before = ps.copy()
    pose = ev.evaluate(ps, np.zeros((0, 3), dtype=np.float32), syn.IDENTITY_TF)
    assert np.array_equal(ps, before) and pose.position == (0.0, 0.0, 0.0)
    # all particles far outside with a_range = 0: "No particle is valid!" (cuda_evaluator.cu:366-369)
    e0 = CudaEvaluator(small[1], False, 0.9, 0.0, 0.0, 100.0)
    far = ps.copy()
    far[:, :3] += 400.0
    with pytest.raises(RuntimeError, match="No particle is valid!"):
        e0.evaluate(far, pts, syn.IDENTITY_TF)
    # resample before any update -> state error; too small a capacity -> capacity error
    lib = capi.load_library()
    with pytest.raises(capi.TsdflocError) as ei:
        e0.resample_systematic(0.0, capacity=16)
    assert ei.value.status == capi.E_STATE
    ev.evaluate(ps, pts, syn.IDENTITY_TF)
    with pytest.raises(capi.TsdflocError) as ei:
        ev.resample_systematic(0.01, capacity=2)
    assert ei.value.status == capi.E_CAPACITY
    out = ev.resample_systematic(0.01, capacity=64)
    assert len(out) == 8
    # bad arguments
    assert lib.tsdfloc_sensor_update(ev.ctx, None, 4, None, 4, None, None) == capi.E_BAD_ARG
    e0.close()


def test_two_contexts_are_independent(oracle, omap, small):
    """No process-wide globals (the reference keeps its map pointers in file-scope device globals, cuda_data.h:17-26)."""
    a = CudaEvaluator(small[1])
    spec2 = syn.box_room_map(likelihood_value, likelihood_init(0.1), resolution=0.1, **ROOM)
    m2 = CudaSubVoxelMap(*spec2.min, *spec2.max, spec2.resolution, spec2.init_value)
    m2.setData(spec2.cells)
    b = CudaEvaluator(m2)
    ps = syn.tracking_particles(40, GT, sigma_xy=0.1)
    pts = _scan(1000)
    ra = oracle.evaluate(omap, common.DEFAULT_PARAMS, ps, pts, syn.IDENTITY_TF)
    rb = oracle.evaluate(common.oracle_map_of(oracle, m2), common.DEFAULT_PARAMS, ps, pts, syn.IDENTITY_TF)
    for _ in range(2):
        _, _, raw_b = b.debug_eval(ps, pts, syn.IDENTITY_TF, want_idx=False)
        _, _, raw_a = a.debug_eval(ps, pts, syn.IDENTITY_TF, want_idx=False)
        assert raw_a.tobytes() == ra["raw"].tobytes() and raw_b.tobytes() == rb["raw"].tobytes()
    a.close()
    b.close()


def test_resampler_edge_cases(oracle, ev):
    rs = SystematicResampler(ev, seed=3)
    # a single particle, all weight on one particle, zeros in between, and u0 at both ends of [0, 1/N)
    for n, hot in [(1, 0), (5, 4), (64, 17), (1000, 999)]:
        ps = np.zeros((n, 7), dtype=np.float32)
        ps[:, 0] = np.arange(n)
        ps[hot, 6] = 1.0
        for u0 in (0.0, float(np.nextafter(np.float32(1.0 / n), np.float32(0)))):
            out, parents = rs.resample(ps, u0=u0, want_parents=True)
            m_ref, p_ref = oracle.systematic_resample(ps[:, 6], u0)
            assert len(out) == m_ref and np.array_equal(parents, p_ref)
    # all-zero weights: the reference loop never draws -> empty cloud
    ps = np.zeros((10, 7), dtype=np.float32)
    out = rs.resample(ps, u0=0.01)
    assert len(out) == 0
    # un-normalised weights (sum 3): the reference recurrence simply keeps drawing -> ~3N particles
    ps = np.zeros((200, 7), dtype=np.float32)
    ps[:, 6] = 3.0 / 200
    with pytest.raises(capi.TsdflocError):
        rs.resample(ps, u0=0.001)          # exceeds the default capacity: reported, not truncated


def test_spatial_evaluation_order_changes_nothing():
    """Spatial order on (counting sort of the particles by map cell in front of the evaluation, tsdfloc_sort.cuh) against
    off: raw weights, normalised weights, mean pose and resampled particles must be byte-identical — including
    particles outside the map and NaN poses, whose cell key is clamped."""
    import common
    from tsdf_localization_b200 import CudaEvaluator, SystematicResampler, synthetic as syn
    _, m = common.box_room()
    pts, _ = syn.make_scan("vlp16", syn.GT_POSE, n_points=3000)
    ps = syn.uniform_particles(20011, (-14.0, -14.0, -1.0), (14.0, 14.0, 7.0))      # partly outside the 25.6 m box
    ps[17, 0] = np.nan
    ps[18, 1] = 1e30
    results = []
    for mode in (0, 1):
        ev = CudaEvaluator(m)
        ev.tune(capi.TUNE_SPATIAL_ORDER, mode)
        mine = ps.copy()
        _, _, raw = ev.debug_eval(mine, pts, syn.CALIB_TF, want_idx=False)
        pose = ev.evaluate(mine, pts, syn.CALIB_TF)
        out = SystematicResampler(ev).resample_resident(len(ps), u0=0.37 / len(ps))
        results.append((raw.tobytes(), mine.tobytes(), np.asarray(pose.position, dtype=np.float64).tobytes(), out.tobytes(), ev.kernel_launches()))
        ev.close()
    assert results[0][:4] == results[1][:4]
    assert results[1][4] > results[0][4]            # the sort kernels really ran


def test_gpu_map_ingest_equals_reference_createTSDFMap():
    """tsdfloc_map_from_chunks_gpu against the UNMODIFIED reference createTSDFMap (map_util.h:17-154, compiled into oracle/_ref
    over an in-memory HighFive stand-in) DIRECTLY: geometry, brick table, voxels and free-space points byte-identical, incl.
    negative chunk coordinates and a chunk order that differs from the name order. The host ingest must agree as well."""
    from oracle_lib import Ref, ref_available
    from test_map_ingest import assert_same_map, synthetic_chunks
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    ref = Ref()
    for chunk_pos, centre in (([(0, 0, 0)], (1000.0, 2000.0, 500.0)),
                              ([(-1, 0, 0), (0, 0, -1), (0, 0, 0), (-1, -1, -1), (1, 0, 0)], (1000.0, 2000.0, 500.0)),
                              ([(2, 1, 0), (10, 1, 0), (3, 1, 0)], (9000.0, 6000.0, 2000.0))):
        data = synthetic_chunks(chunk_pos, seed=len(chunk_pos), centre_mm=centre)
        h, free_ref = ref.create_tsdf_map(chunk_pos, data, 0.1)
        gpu = CudaSubVoxelMap.from_chunks(chunk_pos, data, 0.1, device=0)
        assert_same_map(ref, h, gpu)
        assert gpu.rawData().size > 0 and len(free_ref) > 0
        assert gpu.free_map().tobytes() == free_ref.tobytes()
        host = CudaSubVoxelMap.from_chunks(chunk_pos, data, 0.1)
        assert bytes(host.coef()) == bytes(gpu.coef()) and host.rawData().tobytes() == gpu.rawData().tobytes()
        ref.map_destroy(h)
    with pytest.raises(ValueError):
        CudaSubVoxelMap.from_chunks([(0, 0, 0), (0, 0, 0)], np.zeros((2, 64 ** 3), dtype=np.uint32), device=0)


def test_device_resident_ingest_feeds_the_evaluator(oracle):
    """tsdfloc_create_from_chunks: chunks -> bricks -> evaluation context on the device, the voxels never visiting the host.
    Checked against the reference's createTSDFMap arrays: an evaluator built from THOSE arrays (tsdfloc_create) and the oracle
    on those arrays give the same flat indices, hit counts and raw weights, bit for bit; the free-space points kept on the
    device initialise the same particles as the reference's list handed in from the host."""
    from oracle_lib import Ref, ref_available
    from test_map_ingest import synthetic_chunks
    from tsdf_localization_b200 import ParticleCloud
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    ref = Ref()
    chunk_pos = [(-1, 0, 0), (0, 0, -1), (0, 0, 0), (-1, -1, -1), (1, 0, 0)]
    data = synthetic_chunks(chunk_pos, seed=5)
    h, free_ref = ref.create_tsdf_map(chunk_pos, data, 0.1)
    coef, occ, vox = ref.map_arrays(h)
    resident = CudaEvaluator.from_chunks(chunk_pos, data, 0.1)
    uploaded = CudaEvaluator(CudaSubVoxelMap.from_arrays(coef, occ, vox))
    assert bytes(resident.map_desc()) == bytes(uploaded.map_desc())
    assert resident.free_map_size() == len(free_ref) and uploaded.free_map_size() == 0
    om = oracle.map_from_arrays(coef, occ, vox)
    rng = np.random.default_rng(2)
    # a "scan" of points on the sphere the chunks encode (radius 2.5 m around (1, 2, 0.5)), seen from poses near its centre
    d = rng.normal(size=(3000, 3))
    pts = (2.5 * d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    ps = np.zeros((96, 7), dtype=np.float32)
    ps[:, :3] = np.array([1.0, 2.0, 0.5]) + rng.normal(scale=0.05, size=(96, 3))
    ps[:, 3:6] = rng.normal(scale=0.05, size=(96, 3))
    want = oracle.evaluate(om, common.DEFAULT_PARAMS, ps, pts, syn.IDENTITY_TF, want_idx=True)
    assert want["hits"].sum() > 0.5 * want["idx"].size
    for e in (resident, uploaded):
        idx, hits, raw = e.debug_eval(ps, pts, syn.IDENTITY_TF)
        assert np.array_equal(idx, want["idx"]) and np.array_equal(hits, want["hits"]) and raw.tobytes() == want["raw"].tobytes()
    a = ParticleCloud(resident, seed=9).initialize(5000, (0, 0, 0, 0, 0, 0), (0, 0, 0, 0.1, 0.1, 3.0), mode=capi.INIT_FREE_MAP)
    b = ParticleCloud(uploaded, seed=9).initialize(5000, (0, 0, 0, 0, 0, 0), (0, 0, 0, 0.1, 0.1, 3.0), mode=capi.INIT_FREE_MAP, free_map=free_ref)
    assert a.tobytes() == b.tobytes()
    resident.close()
    uploaded.close()
    ref.map_destroy(h)


@pytest.mark.parametrize("n", [1000, 70001, 1 << 20])
def test_cdf_with_rounding_additions_follows_the_reference_order(oracle, ev, n):
    """Weights spanning 80 binades: the fp64 running sum `s += particle.second` (novel_resampling.h:57) rounds, so its value
    depends on the order of the additions. The parallel scan notices (exactness check) and the CDF is redone in the
    reference's serial order: parents identical to the oracle's serial loop; leading zeros, ties and huge jumps included."""
    rng = np.random.default_rng(n)
    w = 10.0 ** rng.uniform(-24.0, 0.0, n)
    w[: n // 7] = 0.0                                # leading run of zero weights
    w[n // 2] = 0.0
    w = (w / w.sum()).astype(np.float32)
    w[n // 3] = np.float32(2.0 ** -30)               # exact powers of two produce rounding ties
    w[n // 3 + 1] = np.float32(2.0 ** -31)
    ps = np.zeros((n, 7), dtype=np.float32)
    ps[:, 0] = np.arange(n)
    ps[:, 6] = w
    for u0 in (0.0, 0.37 / n):
        out, parents = SystematicResampler(ev).resample(ps, u0=u0, want_parents=True)
        m_ref, parents_ref = oracle.systematic_resample(w, u0)
        assert len(out) == m_ref and np.array_equal(parents, parents_ref)
    assert capi.load_library().tsdfloc_last_cdf_was_exact(ev.ctx) == 0, "the weights were meant to make the parallel scan round"
    ok = np.full(n, 1.0 / n, dtype=np.float32) if n & (n - 1) == 0 else None
    if ok is not None:
        ps[:, 6] = ok
        SystematicResampler(ev).resample(ps, u0=0.0)
        assert capi.load_library().tsdfloc_last_cdf_was_exact(ev.ctx) == 1


def test_stage_times_mirror_the_reference_tasks(ev):
    """tsdfloc_stage_times: init_kernel / exec_kernel / weight_update (cuda_evaluator.cu:127,299,362) + resampling."""
    lib = capi.load_library()
    ps = syn.tracking_particles(2000, GT, sigma_xy=0.2)
    pts = _scan(8000)
    ms = (C.c_float * 4)()
    assert lib.tsdfloc_stage_times(ev.ctx, ms) == capi.E_STATE          # timers are off by default
    ev.tune(capi.TUNE_STAGE_TIMERS, 1)
    try:
        mine = ps.copy()
        ev.evaluate(mine, pts, syn.IDENTITY_TF)
        ev.resample_systematic(0.1 / len(ps), capacity=len(ps) + 400)
        capi.check(lib, ev.ctx, lib.tsdfloc_stage_times(ev.ctx, ms))
        init, exe, upd, res = list(ms)
        assert 0 < init < 5 and 0 < exe < 50 and 0 < upd < 5 and 0 < res < 5
        assert exe > init and exe > upd
    finally:
        ev.tune(capi.TUNE_STAGE_TIMERS, 0)

