"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): flat voxel indices and per-particle hit counts bit-exact; per-particle weights within
1e-5 relative (tolerance of the north star; the un-normalised weights are in fact BIT-EXACT, because the kernel
reproduces the reference's fp32 sequential summation order — a tree sum would differ by up to 1e-3 at P >= 30k);
resampled parents identical given the same weights and U0.
"""
import numpy as np
import pytest

import common
from oracle_lib import NEG_AS_MISS, NEG_REF_DEVICE_SAT
from tsdf_localization_b200 import CudaEvaluator, capi, synthetic as syn

pytestmark = pytest.mark.gpu

WEIGHT_RTOL = 1e-5


@pytest.fixture(scope="module")
def room():
    spec, m = common.box_room()
    return spec, m


@pytest.fixture(scope="module")
def evaluator(room):
    ev = CudaEvaluator(room[1])
    yield ev
    ev.close()


@pytest.fixture(scope="module")
def omap(oracle, room):
    return common.oracle_map_of(oracle, room[1])


def _weights64(oracle, omap, particles, points, tf):
    mats = oracle.pose_matrices(particles, tf)
    w32 = np.empty(len(particles))
    w64 = np.empty(len(particles))
    for i in range(len(particles)):
        w32[i], w64[i] = oracle.pose_weight64(omap, common.DEFAULT_PARAMS, mats[i], points)
    return w32, w64


@pytest.mark.parametrize("tf", [syn.IDENTITY_TF, syn.CALIB_TF], ids=["identity_tf", "calib_tf"])
def test_c1_indices_hits_weights(oracle, omap, evaluator, tf):
    ps, pts, _ = common.config_c1()
    ref = oracle.evaluate(omap, common.DEFAULT_PARAMS, ps, pts, tf, mode=NEG_AS_MISS, want_idx=True)
    assert ref["status"] == 0
    idx, hits, raw = evaluator.debug_eval(ps, pts, tf)
    assert idx.shape == ref["idx"].shape
    mism = int((idx != ref["idx"]).sum())
    assert mism == 0, f"{mism} of {idx.size} flat voxel indices differ"
    assert np.array_equal(hits, ref["hits"])
    assert hits.sum() > 0.3 * idx.size, "workload is degenerate: almost no hits"
    # un-normalised weights: the reference's fp32 sequential sum, bit for bit
    assert np.array_equal(raw, ref["raw"]), f"raw weights differ: max rel {common.rel_err(raw, ref['raw']).max():.3e}"
    # full update through the reference-facing call
    mine = ps.copy()
    pose = evaluator.evaluate(mine, pts, tf)
    assert np.array_equal(mine[:, :6], ps[:, :6]), "poses must not be modified"
    werr = common.rel_err(mine[:, 6], ref["particles"][:, 6])
    assert werr.max() <= WEIGHT_RTOL, f"normalised weight error {werr.max():.3e}"
    assert abs(float(mine[:, 6].astype(np.float64).sum()) - 1.0) < 1e-5
    assert np.allclose(pose.position, ref["mean"][:3], atol=1e-4)
    assert np.allclose(pose.rpy, ref["mean"][3:], atol=1e-4)


def test_c1_reference_gpu_semantics_all_particles(oracle, omap, room):
    """C1 with neg_policy = SATURATE_LIKE_REF_GPU: all 500 particles — the 10 that have a lookup below map.min included
    (tests/golden: negband_particles) — match the reference CUDA evaluator's device semantics (oracle mode
    NEG_REF_DEVICE_SAT, cuda_eval_particles.h:12-67) bit for bit: 512,000 flat indices, hit counts, raw weights; normalised
    weights within the north star's 1e-5. (On this map the saturated lookups land in unallocated border cells, so the default
    MISS policy produces the same bits; tests/test_gpu_edges.py has the map where the two policies differ on 25 % of the pairs.)"""
    from pathlib import Path
    g = np.load(Path(__file__).resolve().parent / "golden" / "c1_reference.npz")
    ps, pts, _ = common.config_c1()
    assert ps.tobytes() == g["particles"].tobytes() and int(g["negband_particles_identity"].sum()) >= 10
    for tf in (syn.IDENTITY_TF, syn.CALIB_TF):
        ref = oracle.evaluate(omap, common.DEFAULT_PARAMS, ps, pts, tf, mode=NEG_REF_DEVICE_SAT, want_idx=True)
        ev = CudaEvaluator(room[1], neg_policy=capi.NEG_SATURATE_LIKE_REF_GPU)
        idx, hits, raw = ev.debug_eval(ps, pts, tf)
        assert np.array_equal(idx, ref["idx"]) and np.array_equal(hits, ref["hits"])
        assert raw.tobytes() == ref["raw"].tobytes()
        mine = ps.copy()
        ev.evaluate(mine, pts, tf)
        assert common.rel_err(mine[:, 6], ref["particles"][:, 6]).max() <= WEIGHT_RTOL
        ev.close()


def test_c1_resample_parents_identical(oracle, evaluator):
    ps, pts, _ = common.config_c1()
    mine = ps.copy()
    evaluator.evaluate(mine, pts, syn.IDENTITY_TF)
    n = len(mine)
    for u0 in (0.0, 0.37 / n, float(np.nextafter(np.float32(1.0 / n), np.float32(0)))):
        out, parents = evaluator.resample_systematic(u0, capacity=n + n // 8 + 64, want_parents=True)
        m_ref, parents_ref = oracle.systematic_resample(mine[:, 6], u0)
        assert len(out) == m_ref
        assert np.array_equal(parents, parents_ref)
        assert np.array_equal(out, mine[parents_ref])


def test_c2_hits_and_weights(oracle, omap, evaluator):
    ps, pts, _ = common.config_c2(2048)
    ref = oracle.evaluate(omap, common.DEFAULT_PARAMS, ps, pts, syn.IDENTITY_TF, mode=NEG_AS_MISS)
    _, hits, raw = evaluator.debug_eval(ps, pts, syn.IDENTITY_TF, want_idx=False)
    assert np.array_equal(hits, ref["hits"])
    assert np.array_equal(raw, ref["raw"]), f"raw weights differ: max rel {common.rel_err(raw, ref['raw']).max():.3e}"
    w32, w64 = _weights64(oracle, omap, ps[:32], pts, syn.IDENTITY_TF)
    print(f"C2: reference fp32-sequential sum vs fp64 sum: {common.rel_err(w32, w64).max():.2e} (what a tree sum would miss)")
    mine = ps.copy()
    evaluator.evaluate(mine, pts, syn.IDENTITY_TF)
    assert common.rel_err(mine[:, 6], ref["particles"][:, 6]).max() <= WEIGHT_RTOL


def test_c1_both_register_budgets_bit_exact(oracle, omap, evaluator):
    """The kernel ships at two register budgets (64: 32 CTAs per SM; 128: 16 CTAs per SM, chosen for slices of at most 16,384
    particles): same code, same bits — indices, hit counts, raw weights against the oracle for both, both pairings."""
    ps, pts, _ = common.config_c1()
    ref = oracle.evaluate(omap, common.DEFAULT_PARAMS, ps, pts, syn.CALIB_TF, mode=NEG_AS_MISS, want_idx=True)
    try:
        for registers in (1, 2):
            for pairing in (1, 2):
                evaluator.tune(capi.TUNE_EVAL_REGISTERS, registers)
                evaluator.tune(capi.TUNE_EVAL_PAIRING, pairing)
                idx, hits, raw = evaluator.debug_eval(ps, pts, syn.CALIB_TF)
                assert np.array_equal(idx, ref["idx"]) and np.array_equal(hits, ref["hits"]), (registers, pairing)
                assert raw.tobytes() == ref["raw"].tobytes(), (registers, pairing)
    finally:
        evaluator.tune(capi.TUNE_EVAL_REGISTERS, 0)
        evaluator.tune(capi.TUNE_EVAL_PAIRING, 0)
