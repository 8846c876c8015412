"""The reference's OWN CUDA evaluator (src/cuda/cuda_evaluator.cu + cuda_eval_particles.h, compiled unmodified for sm_100a
into oracle/_ref/libtsdf_ref_cuda.so) running on the same B200, as a second checker beside the CPU oracle:

  * it pins the oracle's restatement of the device semantics (NEG_REF_DEVICE_SAT): on a map where a quarter of the lookups
    fall below map.min, the product in SATURATE_LIKE_REF_GPU mode agrees with the reference's GPU weights, the default MISS
    policy does not;
  * tolerance, not bit-exactness: the reference's GPU build differs from its CPU build by sinf/cosf (vs fp64 sin/cos) and by
    nvcc's default FMA contraction in the point transform, which moves a few lookups across voxel faces (weights 1e-4 rel.).
"""
import numpy as np
import pytest

import common
from oracle_lib import Ref, ref_cuda_path
from tsdf_localization_b200 import CudaEvaluator, CudaSubVoxelMap, capi, likelihood_init, likelihood_value, synthetic as syn

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_cuda_path().exists(), reason="oracle/_ref/libtsdf_ref_cuda.so not built")]

GT = (0.4, -0.3, 1.2, 0.01, -0.02, 0.4)
ROOM = dict(room_lo=(-3.0, -2.5, 0.0), room_hi=(3.0, 2.5, 3.0))


@pytest.fixture(scope="module")
def refcuda():
    return Ref(cuda=True)


def _ref_gpu_weights(refcuda, spec, ps, pts, tf):
    rm = refcuda.map_create(spec.min, spec.max, spec.resolution, spec.init_value)
    assert refcuda.map_set_data(rm, spec.cells) == 0
    ev = refcuda.eval_create(rm)
    assert ev, refcuda.last_error()
    rc, out, pose, err = refcuda.evaluate(ev, ps, pts, tf, use_cuda=True)
    assert rc == 0, err
    refcuda.eval_destroy(ev)
    refcuda.map_destroy(rm)
    return out[:, 6], pose


def test_reference_cuda_evaluator_agrees_on_c1(refcuda):
    spec, m = common.box_room()
    ps, pts, _ = common.config_c1()
    w_ref, pose_ref = _ref_gpu_weights(refcuda, spec, ps, pts, syn.CALIB_TF)
    ev = CudaEvaluator(m, neg_policy=capi.NEG_SATURATE_LIKE_REF_GPU)
    mine = ps.copy()
    pose = ev.evaluate(mine, pts, syn.CALIB_TF)
    ev.close()
    rel = common.rel_err(mine[:, 6], w_ref)
    print(f"C1 vs the reference CUDA evaluator: max rel {rel.max():.2e}, median {np.median(rel):.2e}")
    assert rel.max() < 2e-3 and np.median(rel) < 1e-5
    np.testing.assert_allclose(pose.position, pose_ref[:3], atol=2e-3)


def test_saturate_policy_is_what_the_reference_gpu_does(refcuda):
    spec = syn.box_room_map(likelihood_value, likelihood_init(0.1), resolution=0.05, margin=0.0, **ROOM)
    m = CudaSubVoxelMap(*spec.min, *spec.max, spec.resolution, spec.init_value)
    m.setData(spec.cells)
    ps = syn.tracking_particles(512, GT, sigma_xy=0.3)
    pts, _ = syn.make_scan("vlp16", GT, n_points=3000, **ROOM)
    w_ref, _ = _ref_gpu_weights(refcuda, spec, ps, pts, syn.IDENTITY_TF)
    got = {}
    for name, policy in (("saturate", capi.NEG_SATURATE_LIKE_REF_GPU), ("miss", capi.NEG_MISS)):
        ev = CudaEvaluator(m, neg_policy=policy)
        mine = ps.copy()
        ev.evaluate(mine, pts, syn.IDENTITY_TF)
        ev.close()
        got[name] = common.rel_err(mine[:, 6], w_ref)
    print(f"negative band: saturate vs reference GPU max rel {got['saturate'].max():.2e}; miss {got['miss'].max():.2e}")
    assert got["saturate"].max() < 5e-3 and np.median(got["saturate"]) < 1e-4
    assert got["miss"].max() > 10 * got["saturate"].max(), "the two policies must be distinguishable on this map"
