"""Runs the reference's OWN benchmark program (src/num_particles_eval.cpp, compiled unmodified into oracle/_ref/num_particles_eval_*
by oracle/Makefile, see oracle/npe_harness.cpp) on a synthetic snapshot: a box room as raw 64^3 TSDF chunks (what createTSDFMap
reads from its HDF5 file) and an .mcl snapshot of a VLP-16 scan inside it (written by tsdfloc_mcl_write).
    python scripts/run_num_particles_eval.py --impl b200,refcuda --num-particles 500000 --inc 100000 --repeat 3
Prints one JSON line per implementation: the program's own "| num. particles | runtime [ms] |" table (integer milliseconds per
evaluate() call, as the reference reports them). b200 = linked against the product's drop-in CudaEvaluator; refcuda = against the
reference's own CUDA evaluator; cpu = use_cuda=false (the reference's OpenMP evaluator). TEST / MEASUREMENT INFRASTRUCTURE."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tsdf_localization_b200 import synthetic as syn                 # noqa: E402
from tsdf_localization_b200.mcl_file import MCLFile                 # noqa: E402

RES = 0.064            # map_util.h: MAP_RESOLUTION of the chunked global map
ROOM_LO, ROOM_HI = np.array([-3.0, -2.5, 0.0]), np.array([3.0, 2.5, 3.0])
GT = (0.4, -0.3, 1.2, 0.01, -0.02, 0.4)


def write_chunk_dump(path):
    """Box room as 64^3 chunks of {int16 tsdf_mm, int16 weight} words (util/tsdf.h:11-87), voxel = corner idx * 0.064 m
    (map_util.h:118-120): every voxel within 1.2 m of a wall carries weight 1 (|tsdf| < 600 mm are stored, the others free space)."""
    band = 1.2
    lo_c = np.floor((ROOM_LO - band) / (64 * RES)).astype(int)
    hi_c = np.floor((ROOM_HI + band) / (64 * RES)).astype(int)
    pos = [(cx, cy, cz) for cx in range(lo_c[0], hi_c[0] + 1) for cy in range(lo_c[1], hi_c[1] + 1) for cz in range(lo_c[2], hi_c[2] + 1)]
    ax = np.arange(64)
    with open(path, "wb") as f:
        f.write(np.int32(len(pos)).tobytes())
        f.write(np.asarray(pos, dtype=np.int32).tobytes())
        for cx, cy, cz in pos:
            p = np.stack(np.meshgrid((64 * cx + ax) * RES, (64 * cy + ax) * RES, (64 * cz + ax) * RES, indexing="ij"), axis=-1)
            d = syn.box_sdf(p, ROOM_LO, ROOM_HI)
            value = np.clip(np.rint(d * 1000.0), -600, 600).astype(np.int16)
            weight = (np.abs(d) < band).astype(np.int16)
            words = value.view(np.uint16).astype(np.uint32) | (weight.view(np.uint16).astype(np.uint32) << 16)
            f.write(np.ascontiguousarray(words).tobytes())       # index 64*64*i + 64*j + k (map_util.h:106)
    return len(pos)


def write_snapshot(path):
    pts, ring = syn.make_scan("vlp16", GT, room_lo=tuple(ROOM_LO), room_hi=tuple(ROOM_HI))
    ps = syn.tracking_particles(16, GT, sigma_xy=0.05, sigma_z=0.05, sigma_yaw=0.03)
    ps[:, 6] = 1.0 / len(ps)
    cr, sr, cp, sp, cy, sy = (f(a / 2) for a in GT[3:] for f in (np.cos, np.sin))
    q = (cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy)   # w x y z
    MCLFile(str(path)).write(pts, ring.astype(np.int32), ps, syn.IDENTITY_TF, GT[0], GT[1], GT[2], *q)
    return len(pts)


def run(impl, mcl, chunks, args, cwd):
    exe = ROOT / "oracle" / "_ref" / {"b200": "num_particles_eval_b200", "refcuda": "num_particles_eval_refcuda", "cpu": "num_particles_eval_cpu"}[impl]
    if not os.access(exe, os.X_OK):
        os.chmod(exe, 0o755)
    env = dict(os.environ, ROSPARAM_num_particles=str(args.num_particles), ROSPARAM_inc=str(args.inc), ROSPARAM_repeat=str(args.repeat),
               ROSPARAM_sigma_trans=str(args.sigma_trans), ROSPARAM_sigma_rot=str(args.sigma_rot), ROSPARAM_use_cuda="0" if impl == "cpu" else "1")
    res = subprocess.run([str(exe), str(mcl), str(chunks)], cwd=cwd, env=env, capture_output=True, text=True, timeout=args.timeout)
    out = res.stdout
    table, seen = [], False
    for line in out.splitlines():
        if line.startswith("| num. particles"):
            seen = True
        elif seen:
            parts = line.split()
            if len(parts) == 2 and parts[0].isdigit():
                table.append([int(parts[0]), int(parts[1])])
            else:
                seen = False
    sizes = [ln.split(":")[1].strip() for ln in out.splitlines() if ln.startswith(("Original cloud size", "Reduced cloud size"))]
    return {"impl": impl, "program": "src/num_particles_eval.cpp (unmodified)", "rc": res.returncode, "scan_points_evaluated": sizes[0] if sizes else None,
            "repeat": args.repeat, "num_particles__runtime_ms": table, "finished": "Evaluation finished!" in out,
            "stderr_tail": res.stderr.strip().splitlines()[-2:] if res.returncode else []}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--num-particles", type=int, default=500000)
    ap.add_argument("--inc", type=int, default=100000)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--sigma-trans", type=float, default=0.05)
    ap.add_argument("--sigma-rot", type=float, default=0.03)
    ap.add_argument("--timeout", type=float, default=600.0)
    args = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        n_chunks = write_chunk_dump(tmp / "room.chunks")
        n_pts = write_snapshot(tmp / "snapshot.mcl")
        for impl in args.impl.split(","):
            r = run(impl, tmp / "snapshot.mcl", tmp / "room.chunks", args, tmp)
            r.update(map_chunks=n_chunks, snapshot_points=n_pts)
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
