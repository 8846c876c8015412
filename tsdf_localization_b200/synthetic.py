"""Synthetic workloads shared by tests/ and bench.py (BASELINE.json configs). Pure numpy, seeds fixed.

Maps are analytic signed-distance fields quantised like the real pipeline (int16 millimetres, |tsdf| < 600 mm kept,
value = N(d; 0, 0.1)^3) and handed to the map builder as (x, y, z, value) cells at voxel CENTRES, so fp32 rounding
cannot move a cell across a voxel face. Scans are ray-cast from a ground-truth pose (ring-major point order, Gaussian
range noise, |p| < 1 m dropped as src/evaluation/tsdf_evaluator.cpp:319 does). Particle clouds mirror
src/num_particles_eval.cpp:232 (tracking) and src/particle_cloud.cpp:104-146 (global).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

SIGMA = 0.1            # likelihood_evaluation.h:21
TRUNCATION_MM = 600    # grid_map.h:24


@dataclass
class MapSpec:
    min: Tuple[float, float, float]
    max: Tuple[float, float, float]
    resolution: float
    cells: np.ndarray          # [n, 4] float32 (x, y, z, value)
    init_value: float


def _likelihood_lut(value_fn) -> np.ndarray:
    """value for tsdf_mm in (-600, 600), indexed by tsdf_mm + 599."""
    return np.array([value_fn(float(mm), SIGMA) for mm in range(-TRUNCATION_MM + 1, TRUNCATION_MM)], dtype=np.float32)


def box_sdf(p: np.ndarray, lo: np.ndarray, hi: np.ndarray) -> np.ndarray:
    """Signed distance to the surface of the axis-aligned box [lo, hi], positive INSIDE (free space)."""
    c = 0.5 * (lo + hi)
    h = 0.5 * (hi - lo)
    q = np.abs(p - c) - h
    outside = np.linalg.norm(np.maximum(q, 0.0), axis=-1)
    inside = np.minimum(q.max(axis=-1), 0.0)
    return -(outside + inside)


CHUNK_VOXELS = 64      # grid_map.h:18 CHUNK_SIZE: the real pipeline's maps are unions of 64^3-voxel HDF5 chunks


def chunk_aligned_bounds(room_lo, room_hi, resolution: float):
    """Bounding box createTSDFMap would compute for this room (map_util.h:27-66): the union of the 64^3-voxel chunks that
    hold any stored voxel, i.e. the room grown by the truncation band and rounded outwards to chunk edges."""
    edge = CHUNK_VOXELS * resolution
    band = TRUNCATION_MM * 1e-3
    lo = np.floor((np.asarray(room_lo, dtype=np.float64) - band) / edge + 1e-9) * edge
    hi = np.ceil((np.asarray(room_hi, dtype=np.float64) + band) / edge - 1e-9) * edge
    return lo, hi


def box_room_map(value_fn, init_value: float, room_lo=(-10.0, -10.0, 0.0), room_hi=(10.0, 10.0, 5.0), resolution: float = 0.05,
                 margin: float | None = None) -> MapSpec:
    """Config C1–C3 map: an empty box room with the +-0.6 m truncation band stored around all six faces.
    margin=None (default): chunk-aligned bounding box like the real pipeline's (20x20x5 m @ 5 cm -> 25.6x25.6x9.6 m box);
    margin=m: bounding box = room grown by m on every side (m = 0 puts the walls ON the box faces, which makes a quarter
    of all lookups land at negative offsets — the reference's undefined-behaviour band, used by the policy tests)."""
    if margin is None:
        lo, hi = chunk_aligned_bounds(room_lo, room_hi, resolution)
    else:
        lo = np.asarray(room_lo, dtype=np.float64) - margin
        hi = np.asarray(room_hi, dtype=np.float64) + margin
    dims = np.ceil((hi - lo) / resolution - 1e-9).astype(np.int64)
    lut = _likelihood_lut(value_fn)
    rlo, rhi = np.asarray(room_lo, dtype=np.float64), np.asarray(room_hi, dtype=np.float64)
    out = []
    xs = lo[0] + (np.arange(dims[0]) + 0.5) * resolution
    ys = lo[1] + (np.arange(dims[1]) + 0.5) * resolution
    for kz in range(int(dims[2])):
        z = lo[2] + (kz + 0.5) * resolution
        X, Y = np.meshgrid(xs, ys, indexing="ij")
        P = np.stack([X, Y, np.full_like(X, z)], axis=-1).reshape(-1, 3)
        d = box_sdf(P, rlo, rhi)
        mm = np.clip(np.rint(d * 1000.0), -32768, 32767).astype(np.int64)
        keep = np.abs(mm) < TRUNCATION_MM
        if not keep.any():
            continue
        vals = lut[mm[keep] + TRUNCATION_MM - 1]
        out.append(np.concatenate([P[keep].astype(np.float32), vals[:, None]], axis=1))
    cells = np.concatenate(out, axis=0).astype(np.float32)
    return MapSpec(tuple(lo.tolist()), tuple(hi.tolist()), resolution, cells, init_value)


# ---- scans ---------------------------------------------------------------------------------------------------------

def rpy_matrix(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """R = Rz(yaw) Ry(pitch) Rx(roll), the reference's Euler convention (tsdf_evaluator.cpp:115-128)."""
    sa, ca, sb, cb, sg, cg = np.sin(roll), np.cos(roll), np.sin(pitch), np.cos(pitch), np.sin(yaw), np.cos(yaw)
    return np.array([[cb * cg, sa * sb * cg - ca * sg, ca * sb * cg + sa * sg],
                     [cb * sg, sa * sb * sg + ca * cg, ca * sb * sg - sa * cg],
                     [-sb, sa * cb, ca * cb]])


def lidar_directions(rings: int, azimuths: int, fov_deg: float) -> Tuple[np.ndarray, np.ndarray]:
    """Unit ray directions in the sensor frame, ring-major; returns (dirs [rings*azimuths, 3], ring index)."""
    elev = np.deg2rad(np.linspace(-fov_deg, fov_deg, rings))
    az = np.linspace(-np.pi, np.pi, azimuths, endpoint=False)
    E, A = np.meshgrid(elev, az, indexing="ij")
    d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], axis=-1).reshape(-1, 3)
    ring = np.repeat(np.arange(rings, dtype=np.int32), azimuths)
    return d, ring


def raycast_box(pose6, dirs: np.ndarray, room_lo, room_hi, noise_sigma: float = 0.01, seed: int = 1) -> np.ndarray:
    """Ranges to the walls of an empty box from inside it, plus Gaussian range noise; points in the SENSOR frame."""
    o = np.asarray(pose6[:3], dtype=np.float64)
    R = rpy_matrix(*pose6[3:6])
    dw = dirs @ R.T
    lo, hi = np.asarray(room_lo, dtype=np.float64), np.asarray(room_hi, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_axis = np.where(dw > 0, (hi - o) / dw, np.where(dw < 0, (lo - o) / dw, np.inf))
    t = t_axis.min(axis=1)
    rng = np.random.default_rng(seed)
    t = t + rng.normal(0.0, noise_sigma, size=t.shape)
    return (dirs * t[:, None]).astype(np.float32)


def make_scan(kind: str, pose6, room_lo=(-10.0, -10.0, 0.0), room_hi=(10.0, 10.0, 5.0), seed: int = 1,
              n_points: int | None = None) -> Tuple[np.ndarray, np.ndarray]:
    """kind: 'os1-128' (128 x 1024 = 131,072 points), 'vlp16' (16 x 1875 = 30,000). Optional uniform sub-sampling to
    n_points (C1: 1,024 of the VLP-16 scan). Returns (points float32[P,3] ring-major, ring int32[P])."""
    if kind == "os1-128":
        dirs, ring = lidar_directions(128, 1024, 22.5)
    elif kind == "vlp16":
        dirs, ring = lidar_directions(16, 1875, 15.0)
    else:
        raise ValueError(kind)
    pts = raycast_box(pose6, dirs, room_lo, room_hi, seed=seed)
    keep = np.linalg.norm(pts.astype(np.float64), axis=1) >= 1.0
    pts, ring = pts[keep], ring[keep]
    if n_points is not None and n_points < pts.shape[0]:
        sel = np.linspace(0, pts.shape[0] - 1, n_points).astype(np.int64)
        pts, ring = pts[sel], ring[sel]
    return np.ascontiguousarray(pts), np.ascontiguousarray(ring)


# ---- particles -----------------------------------------------------------------------------------------------------

def tracking_particles(n: int, gt_pose6, seed: int = 42, sigma_xy: float = 0.5, sigma_z: float = 0.1, sigma_rp: float = 0.02,
                       sigma_yaw: float = 0.5) -> np.ndarray:
    """N(gt, sigma) cloud, weight slot 0 (num_particles_eval.cpp:232 operating point)."""
    rng = np.random.default_rng(seed)
    p = np.zeros((n, 7), dtype=np.float32)
    gt = np.asarray(gt_pose6, dtype=np.float64)
    sig = np.array([sigma_xy, sigma_xy, sigma_z, sigma_rp, sigma_rp, sigma_yaw])
    p[:, :6] = (gt[None, :] + rng.normal(size=(n, 6)) * sig[None, :]).astype(np.float32)
    return p


def uniform_particles(n: int, lo, hi, seed: int = 7) -> np.ndarray:
    """Global localisation: xyz uniform in [lo, hi], roll/pitch/yaw uniform in [-pi, pi] (particle_cloud.cpp:62-101)."""
    rng = np.random.default_rng(seed)
    p = np.zeros((n, 7), dtype=np.float32)
    p[:, :3] = rng.uniform(np.asarray(lo), np.asarray(hi), size=(n, 3)).astype(np.float32)
    p[:, 3:6] = rng.uniform(-np.pi, np.pi, size=(n, 3)).astype(np.float32)
    return p


GT_POSE = (1.3, -2.1, 1.5, 0.01, -0.02, 0.4)   # ground-truth sensor pose inside the box room
IDENTITY_TF = np.eye(4, dtype=np.float32).reshape(-1)
# a non-trivial scanner->base calibration: 10 cm forward, 30 cm up, 2 degrees of pitch
CALIB_TF = np.array([0.99939083, 0.0, 0.0348995, 0.10,
                     0.0, 1.0, 0.0, 0.0,
                     -0.0348995, 0.0, 0.99939083, 0.30,
                     0.0, 0.0, 0.0, 1.0], dtype=np.float32)
