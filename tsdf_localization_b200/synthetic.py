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


# ---- large sparse maps (configs C4 / C5) ---------------------------------------------------------------------------------

def grid_rooms_arrays(make_map, value_fn, init_value: float, extent=(100.0, 100.0, 10.0), room: float = 20.0,
                      resolution: float = 0.05, ceiling: bool = False):
    """Multi-room building for the global-localisation / larger-than-L2 configs: walls on every `room`-metre grid line in x
    and y up to height extent[2], a floor at z = 0 (optionally a ceiling), stored as the +-0.6 m truncation band only.

    Builds the reference's two arrays DIRECTLY (bricks in increasing upper-cell order, unset voxels = init_value; the layout
    CudaSubVoxelMap::setData produces, cuda_sub_voxel_map.tcc:170-230) instead of going through a cell list, because the
    full-size map has ~2e8 stored voxels. `make_map(min, max, res, init)` must return an empty product CudaSubVoxelMap (its
    coef() supplies the geometry). Returns (desc, grid_occ int32[], data float32[]). tests/test_synthetic.py checks a small
    instance against setData on the equivalent cell list."""
    ex = np.asarray(extent, dtype=np.float64)
    lo, hi = chunk_aligned_bounds((0.0, 0.0, 0.0), ex, resolution)
    m = make_map(tuple(lo.tolist()), tuple(hi.tolist()), resolution, init_value)
    c = m.coef()
    sub = int(c.sub_dim)
    up = [int(c.up_dim[a]) for a in range(3)]
    mn = [float(np.float32(c.min[a])) for a in range(3)]
    lut = _likelihood_lut(value_fn)
    band = TRUNCATION_MM * 1e-3

    def axis_coords(a):
        k = np.arange(up[a] * sub, dtype=np.float64)
        return mn[a] + (k + 0.5) * resolution            # voxel centres

    X, Y, Z = axis_coords(0), axis_coords(1), axis_coords(2)
    dx = np.abs(X - np.clip(np.rint(X / room), 0, round(ex[0] / room)) * room)
    dy = np.abs(Y - np.clip(np.rint(Y / room), 0, round(ex[1] / room)) * room)
    dz = np.abs(Z)
    if ceiling:
        dz = np.minimum(dz, np.abs(Z - ex[2]))
    # walls exist for 0 <= z <= H and inside the footprint (+ band); beyond that only the floor/ceiling planes count
    wall_z = (Z >= -band) & (Z <= ex[2] + band)
    in_x = (X >= -band) & (X <= ex[0] + band)
    in_y = (Y >= -band) & (Y <= ex[1] + band)
    big = 1e9
    dxw = np.where(in_x, dx, big)
    dyw = np.where(in_y, dy, big)

    def near(d1):      # per upper cell: does any voxel column of it come within the band
        return (d1.reshape(-1, sub) < band).any(axis=1)

    nx, ny, nz = near(dxw), near(dyw), near(np.where(wall_z, dz, big) if not ceiling else dz)
    zw = wall_z.reshape(-1, sub).any(axis=1)
    fx, fy = in_x.reshape(-1, sub).any(axis=1), in_y.reshape(-1, sub).any(axis=1)
    cand = ((nx[:, None, None] & fy[None, :, None] & zw[None, None, :]) | (fx[:, None, None] & ny[None, :, None] & zw[None, None, :])
            | (fx[:, None, None] & fy[None, :, None] & nz[None, None, :]))
    uxs, uys, uzs = np.nonzero(cand)
    order = np.argsort(uxs + uys * up[0] + uzs * up[0] * up[1], kind="stable")
    uxs, uys, uzs = uxs[order], uys[order], uzs[order]
    brick = sub ** 3
    grid_occ = np.full(up[0] * up[1] * up[2], -1, dtype=np.int32)
    data = np.full(len(uxs) * brick, np.float32(init_value), dtype=np.float32)
    n_alloc = 0
    for ux, uy, uz in zip(uxs.tolist(), uys.tolist(), uzs.tolist()):
        sx, sy, sz = slice(ux * sub, (ux + 1) * sub), slice(uy * sub, (uy + 1) * sub), slice(uz * sub, (uz + 1) * sub)
        wz = wall_z[sz][:, None, None]
        d_wall = np.minimum(np.where(in_y[sy][None, :, None], dxw[sx][None, None, :], big),
                            np.where(in_x[sx][None, None, :], dyw[sy][None, :, None], big))
        d_floor = np.where(in_x[sx][None, None, :] & in_y[sy][None, :, None], dz[sz][:, None, None], big)
        d = np.minimum(np.where(wz, d_wall, big), d_floor)          # [sz][sy][sx]
        mm = np.rint(d * 1000.0)
        keep = mm < TRUNCATION_MM
        if not keep.any():
            continue
        vals = np.full(d.shape, np.float32(init_value), dtype=np.float32)
        vals[keep] = lut[mm[keep].astype(np.int64) + TRUNCATION_MM - 1]
        off = n_alloc * brick
        data[off:off + brick] = vals.reshape(-1)
        grid_occ[ux + uy * up[0] + uz * up[0] * up[1]] = off
        n_alloc += 1
    data = data[:n_alloc * brick].copy()
    import copy
    desc = copy.copy(c)
    desc = type(c).from_buffer_copy(bytes(c))
    desc.data_size = n_alloc * brick
    return desc, grid_occ, data


def grid_rooms_cells(value_fn, extent, room: float, resolution: float, ceiling: bool = False) -> np.ndarray:
    """The same building as an explicit (x, y, z, value) cell list at voxel centres (small instances only: the check that
    grid_rooms_arrays equals what setData builds)."""
    ex = np.asarray(extent, dtype=np.float64)
    lo, hi = chunk_aligned_bounds((0.0, 0.0, 0.0), ex, resolution)
    band = TRUNCATION_MM * 1e-3
    lut = _likelihood_lut(value_fn)
    flo = np.float32(lo).astype(np.float64)
    n = np.ceil((hi - lo) / resolution - 1e-9).astype(np.int64)
    X = flo[0] + (np.arange(n[0]) + 0.5) * resolution
    Y = flo[1] + (np.arange(n[1]) + 0.5) * resolution
    Z = flo[2] + (np.arange(n[2]) + 0.5) * resolution
    big = 1e9
    in_x, in_y = (X >= -band) & (X <= ex[0] + band), (Y >= -band) & (Y <= ex[1] + band)
    wall_z = (Z >= -band) & (Z <= ex[2] + band)
    dx = np.where(in_x, np.abs(X - np.clip(np.rint(X / room), 0, round(ex[0] / room)) * room), big)
    dy = np.where(in_y, np.abs(Y - np.clip(np.rint(Y / room), 0, round(ex[1] / room)) * room), big)
    dz = np.abs(Z)
    if ceiling:
        dz = np.minimum(dz, np.abs(Z - ex[2]))
    out = []
    for kz in range(len(Z)):
        d_wall = np.minimum(np.where(in_y[None, :], dx[:, None], big), np.where(in_x[:, None], dy[None, :], big))
        d_floor = np.where(in_x[:, None] & in_y[None, :], dz[kz], big)
        d = np.minimum(d_wall if wall_z[kz] else big, d_floor)
        mm = np.rint(d * 1000.0)
        keep = mm < TRUNCATION_MM
        if not keep.any():
            continue
        ix, iy = np.nonzero(keep)
        vals = lut[mm[keep].astype(np.int64) + TRUNCATION_MM - 1]
        out.append(np.stack([X[ix], Y[iy], np.full(len(ix), Z[kz]), vals], axis=1).astype(np.float32))
    return np.concatenate(out, axis=0)


def raycast_rooms(pose6, dirs: np.ndarray, room: float, height: float, noise_sigma: float = 0.01, seed: int = 1) -> np.ndarray:
    """Scan inside one room of the grid building (no ceiling: rays that leave above the walls give no return)."""
    o = np.asarray(pose6[:3], dtype=np.float64)
    lo = np.array([np.floor(o[0] / room) * room, np.floor(o[1] / room) * room, 0.0])
    hi = np.array([lo[0] + room, lo[1] + room, 1e9])
    pts = raycast_box(pose6, dirs, lo, hi, noise_sigma=noise_sigma, seed=seed)
    R = rpy_matrix(*pose6[3:6])
    zw = (pts.astype(np.float64) @ R.T)[:, 2] + o[2]
    ok = np.isfinite(pts).all(axis=1) & (zw <= height) & (np.linalg.norm(pts.astype(np.float64), axis=1) < 200.0)
    return pts, ok


def reduce_scan(points: np.ndarray, ring: np.ndarray, cell: float) -> Tuple[np.ndarray, np.ndarray]:
    """Ring-aware voxel reduction of TSDFEvaluator::evaluateParticles (src/evaluation/tsdf_evaluator.cpp:304-376) for clouds
    WITHOUT points nearer than 1 m (where the reference's ring iterator desynchronises): per (ring, cell) the first point in
    scan order is kept; output ordered by (ring, original index). Returns (points, ring)."""
    p = np.ascontiguousarray(points, dtype=np.float32)
    res = np.float32(cell)
    key = np.floor(p / res).astype(np.int64)          # float32 division + floor, as the reference evaluates it
    full = np.concatenate([ring.astype(np.int64)[:, None], key], axis=1)
    _, first = np.unique(full, axis=0, return_index=True)
    first = np.sort(first)
    order = np.lexsort((first, ring[first]))
    sel = first[order]
    return p[sel], ring[sel]
