"""Workload builders shared by the tests: BASELINE.json configs through the product's own host map builder."""
from __future__ import annotations

import functools

import numpy as np

from tsdf_localization_b200 import CudaSubVoxelMap, likelihood_init, likelihood_value
from tsdf_localization_b200 import synthetic as syn

DEFAULT_PARAMS = (0.9, 0.1, 0.0, 100.0)   # a_hit, a_range, a_max, max_range (util.h:13-18)


@functools.lru_cache(maxsize=4)
def box_room(resolution: float = 0.05, margin=None, small: bool = False):
    """(spec, product map). small=True: a 6x5x3 m room for fast CPU tests."""
    if small:
        spec = syn.box_room_map(likelihood_value, likelihood_init(syn.SIGMA), room_lo=(-3.0, -2.5, 0.0), room_hi=(3.0, 2.5, 3.0),
                                resolution=resolution, margin=margin)
    else:
        spec = syn.box_room_map(likelihood_value, likelihood_init(syn.SIGMA), resolution=resolution, margin=margin)
    m = CudaSubVoxelMap(*spec.min, *spec.max, spec.resolution, spec.init_value)
    m.setData(spec.cells)
    return spec, m


def oracle_map_of(oracle, m: CudaSubVoxelMap):
    """Oracle-side map adopting the product map's arrays (same bytes on both sides)."""
    return oracle.map_from_arrays(m.coef(), m.rawGridOcc(), m.rawData())


def config_c1():
    """C1: 500 particles x 1,024 points sub-sampled from a VLP-16 scan, box room 5 cm."""
    pts, ring = syn.make_scan("vlp16", syn.GT_POSE, n_points=1024)
    ps = syn.tracking_particles(500, syn.GT_POSE)
    return ps, pts, ring


def config_c2(n_particles: int = 5000):
    pts, ring = syn.make_scan("vlp16", syn.GT_POSE)
    ps = syn.tracking_particles(n_particles, syn.GT_POSE)
    return ps, pts, ring


def config_c3(n_particles: int = 65536):
    pts, ring = syn.make_scan("os1-128", syn.GT_POSE)
    ps = syn.tracking_particles(n_particles, syn.GT_POSE)
    return ps, pts, ring


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


@functools.lru_cache(maxsize=2)
def grid_rooms(extent=(100.0, 100.0, 10.0), room: float = 20.0, resolution: float = 0.05):
    """C4/C5 map (1.4 GB at the default size; exceeds the 126 MB L2): product map adopting directly-built arrays."""
    mk = lambda mn, mx, res, init: CudaSubVoxelMap(*mn, *mx, res, init)   # noqa: E731
    desc, occ, data = syn.grid_rooms_arrays(mk, likelihood_value, likelihood_init(syn.SIGMA), extent=extent, room=room,
                                            resolution=resolution)
    return CudaSubVoxelMap.from_arrays(desc, occ, data)


ROOMS_POSE = (47.0, 52.5, 1.5, 0.01, -0.02, 0.4)


def rooms_scan(reduce_cell=None, extent=(100.0, 100.0, 10.0), room: float = 20.0, pose=ROOMS_POSE):
    dirs, ring = syn.lidar_directions(128, 1024, 22.5)
    pts, ok = syn.raycast_rooms(pose, dirs, room, extent[2])
    pts, ring = np.ascontiguousarray(pts[ok]), np.ascontiguousarray(ring[ok])
    if reduce_cell:
        pts, ring = syn.reduce_scan(pts, ring, reduce_cell)
    return np.ascontiguousarray(pts), ring


def config_c4(n_particles: int = 1 << 20):
    """C4: global localisation — particles uniform over the 100x100 m building, OS1-128 scan after the ring-aware 0.256 m
    reduction (launch/mcl_3d_hilti.launch:81)."""
    pts, ring = rooms_scan(reduce_cell=0.256)
    ps = syn.uniform_particles(n_particles, (0.5, 0.5, 0.3), (99.5, 99.5, 2.5))
    return ps, pts, ring


def config_c5(n_particles: int = 262144):
    """C5: tracking cloud inside the larger-than-L2 map, full OS1-128 scan."""
    pts, ring = rooms_scan()
    ps = syn.tracking_particles(n_particles, ROOMS_POSE)
    return ps, pts, ring
