"""``.mcl`` snapshot files — the reference's fixture format for one scan + particle set
(include/tsdf_localization/util/mcl_file.h:21-32, src/util/mcl_file.cpp:14-113; written by snap_shot_node, replayed by
num_particles_eval / cuda_test_eval / snap_vis_node). Whitespace-separated text:

    P                       number of points
    P x "x y z"             points (sensor frame)
    P x "ring"              ring of every point
    N                       number of particles
    N x "x y z roll pitch yaw  weight"
    16 floats               scanner->robot transform, row-major
    "x y z q1 q2 q3 q4"     reference pose

Numbers are formatted like the reference's ``ostream << float`` (``%g``, 6 significant digits) so files written here are
byte-identical to the reference's for the same values; reading accepts anything ``istream >> float`` accepts. The work is
done by the C library (``tsdfloc_mcl_read`` / ``tsdfloc_mcl_write``, include/tsdfloc.h), which the C++ host side binds too.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi


@dataclass
class MCLSnapshot:
    points: np.ndarray        # float32 [P, 3]
    rings: np.ndarray         # int32 [P]
    particles: np.ndarray     # float32 [N, 7]  x y z roll pitch yaw weight
    tf: np.ndarray            # float32 [16]
    pose: np.ndarray          # float32 [7]  x y z q1 q2 q3 q4


class MCLFile:
    """Thin binding of tsdfloc_mcl_read / tsdfloc_mcl_write (csrc/tsdfloc_mcl.inc): the parsing and formatting happen in the
    C library, so the C++ host side and Python read and write the same bytes."""

    def __init__(self, file_name):
        self.name_ = str(file_name)

    def write(self, points, rings, particles, tf, x, y, z, q_1, q_2, q_3, q_4) -> None:
        lib = capi.load_library()
        pts = np.ascontiguousarray(np.asarray(points, dtype=np.float32).reshape(-1, 3))
        rg = np.ascontiguousarray(np.asarray(rings).reshape(-1), dtype=np.int32)
        ps = np.ascontiguousarray(np.asarray(particles, dtype=np.float32).reshape(-1, 7))
        tfm = np.ascontiguousarray(np.asarray(tf, dtype=np.float32).reshape(-1))
        if len(rg) != len(pts) or len(tfm) != 16:
            raise ValueError("rings must match points; tf must hold 16 values")
        pose = np.array([x, y, z, q_1, q_2, q_3, q_4], dtype=np.float32)
        rc = lib.tsdfloc_mcl_write(self.name_.encode(), pts.ctypes.data_as(C.c_void_p), rg.ctypes.data_as(C.c_void_p), len(pts),
                                   ps.ctypes.data_as(C.c_void_p), len(ps), tfm.ctypes.data_as(C.POINTER(C.c_float)),
                                   pose.ctypes.data_as(C.POINTER(C.c_float)))
        if rc != capi.OK:
            raise OSError(lib.tsdfloc_last_error(None).decode())

    def read(self) -> MCLSnapshot:
        lib = capi.load_library()
        h = C.c_void_p()
        rc = lib.tsdfloc_mcl_read(self.name_.encode(), C.byref(h))
        if rc == capi.E_STATE:
            raise OSError(lib.tsdfloc_last_error(None).decode())
        if rc != capi.OK:
            raise ValueError(lib.tsdfloc_last_error(None).decode())
        try:
            n_points, n_particles = int(lib.tsdfloc_mcl_n_points(h)), int(lib.tsdfloc_mcl_n_particles(h))

            def arr(ptr, shape, dtype):
                if int(np.prod(shape)) == 0:
                    return np.zeros(shape, dtype=dtype)
                return np.ctypeslib.as_array(ptr, shape=shape).astype(dtype, copy=True)

            return MCLSnapshot(arr(lib.tsdfloc_mcl_points(h), (n_points, 3), np.float32), arr(lib.tsdfloc_mcl_rings(h), (n_points,), np.int32),
                               arr(lib.tsdfloc_mcl_particles(h), (n_particles, 7), np.float32), arr(lib.tsdfloc_mcl_tf(h), (16,), np.float32),
                               arr(lib.tsdfloc_mcl_pose(h), (7,), np.float32))
        finally:
            lib.tsdfloc_mcl_free(h)
