#!/usr/bin/env python
"""Mints tests/golden/c1_reference.npz from the UNMODIFIED reference compiled here (oracle/_ref/libtsdf_ref.so).

TEST INFRASTRUCTURE. Run in the container that has /root/reference (after `make -C oracle`):
    python oracle/gen_golden.py
The reference ships no golden vectors; these are outputs of its own code on BASELINE config C1
(box room 20x20x5 m @ 5 cm, 500 particles x 1,024 points), so the GPU box — where /root/reference does not exist —
can still check the oracle and the CUDA path against the reference itself.

What comes from where:
  map_sha / occ_sha        CudaSubVoxelMap::setData arrays                       (verbatim reference)
  raw_*                    TSDFEvaluator::evaluatePose per particle, fp32 seq.   (verbatim; matrices = oracle pose_matrix,
                                                                                  itself pinned through evaluate())
  negband_* / *_policy_*   particles with a lookup at a NEGATIVE axis offset (the reference's undefined-behaviour band,
                           SURVEY 2.5(4)) and the weights under the product's documented policy there (miss)   (oracle)
  norm_* / mean_*          TSDFEvaluator::evaluate normalised weights + pose     (verbatim; OpenMP reduction order makes
                                                                                  the divisor vary in the last ulp)
  rs_*                     SystematicResampler::resample, mt19937 seeds 1..3     (verbatim, incl. the U0 it drew)
  idx_sha_* / hits_* / idx_head_*   flat voxel indices + hit counts              (oracle restatement: getIndex is private
                                                                                  in the reference; pinned via getEntry)
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import common  # noqa: E402
from oracle_lib import NEG_AS_MISS, NEG_REF_HOST_X86, Oracle, Ref  # noqa: E402
from tsdf_localization_b200 import synthetic as syn  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    oracle, ref = Oracle(), Ref()
    spec, m = common.box_room()
    rm = ref.map_create(spec.min, spec.max, spec.resolution, spec.init_value)
    assert ref.map_set_data(rm, spec.cells) == 0
    coef, occ, data = ref.map_arrays(rm)
    assert np.array_equal(occ, m.rawGridOcc()) and data.tobytes() == m.rawData().tobytes(), "product map builder != reference"
    om = oracle.map_from_arrays(coef, occ, data)
    ps, pts, ring = common.config_c1()
    g = dict(particles=ps, points=pts, map_sha=sha(data), occ_sha=sha(occ), data_size=np.uint64(coef.data_size),
             n_bricks=np.uint64((occ >= 0).sum()))
    ev = ref.eval_create(rm)
    for name, tf in (("identity", syn.IDENTITY_TF), ("calib", syn.CALIB_TF)):
        mats = oracle.pose_matrices(ps, tf)
        raw = ref.pose_weights(ev, mats, pts)
        rc, ps_ref, pose, err = ref.evaluate(ev, ps, pts, tf)
        assert rc == 0, err
        o = oracle.evaluate(om, common.DEFAULT_PARAMS, ps, pts, tf, mode=NEG_AS_MISS, want_idx=True)
        o_x86 = oracle.evaluate(om, common.DEFAULT_PARAMS, ps, pts, tf, mode=NEG_REF_HOST_X86, want_idx=True)
        # the reference CPU build, restated with its x86 float->unsigned behaviour, must match it bit for bit everywhere
        assert o_x86["raw"].tobytes() == raw.tobytes(), "oracle (x86 mode) raw weights != verbatim reference"
        np.testing.assert_allclose(o_x86["particles"][:, 6], ps_ref[:, 6], rtol=3e-7)
        # negative band: (particle, point) pairs with a NEGATIVE axis offset, where the reference is undefined behaviour and
        # the product's policy (miss) differs from what this x86 build happens to do. Outside it: identical.
        band = o["idx"] != o_x86["idx"]
        clean = ~band.any(axis=1)
        assert o["raw"][clean].tobytes() == raw[clean].tobytes(), "oracle raw weights != verbatim reference outside the negative band"
        print(f"{name}: negative-band pairs {int(band.sum())} of {band.size} in {int((~clean).sum())} of {len(clean)} particles")
        g[f"negband_particles_{name}"] = ~clean
        g[f"negband_pairs_{name}"] = np.uint64(band.sum())
        g[f"raw_policy_{name}"] = o["raw"]
        g[f"norm_policy_{name}"] = o["particles"][:, 6].copy()
        g[f"tf_{name}"] = np.asarray(tf, dtype=np.float32)
        g[f"raw_{name}"] = raw
        g[f"norm_{name}"] = ps_ref[:, 6].copy()
        g[f"mean_xyz_{name}"] = np.asarray(pose[:3])
        g[f"mean_quat_{name}"] = np.asarray(pose[3:])
        g[f"idx_sha_{name}"] = sha(o["idx"])
        g[f"idx_head_{name}"] = o["idx"][:16].copy()
        g[f"hits_{name}"] = o["hits"]
    # resampling on the reference's own normalised weights (identity tf)
    rc, ps_ref, _, _ = ref.evaluate(ev, ps, pts, syn.IDENTITY_TF)
    g["rs_input"] = ps_ref
    for seed in (1, 2, 3):
        m_out, out, u0 = ref.systematic_resample(ps_ref, seed)
        g[f"rs_u0_{seed}"] = np.float32(u0)
        g[f"rs_out_{seed}"] = out
        mo, parents = oracle.systematic_resample(ps_ref[:, 6], u0)
        assert mo == m_out and np.array_equal(ps_ref[parents], out)
        g[f"rs_parents_{seed}"] = parents
    dst = ROOT / "tests" / "golden" / "c1_reference.npz"
    dst.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(dst, **g)
    print("wrote", dst, dst.stat().st_size, "bytes; threads of the reference build:", ref.omp_threads())


if __name__ == "__main__":
    main()
