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
