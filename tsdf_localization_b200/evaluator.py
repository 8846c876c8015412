"""Host-side mirror of the reference's evaluator / resampler interface, on top of the C ABI (include/tsdfloc.h).

Same names, argument meaning and error behaviour as the reference classes (paths relative to the reference repo):

  CudaSubVoxelMap      include/tsdf_localization/cuda/cuda_sub_voxel_map.h:17-342   (host half: geometry + setData)
  CudaEvaluator        include/tsdf_localization/cuda/cuda_evaluator.h:109-169, src/cuda/cuda_evaluator.cu:21-59,118-428
  TSDFEvaluator        include/tsdf_localization/evaluation/tsdf_evaluator.h:38-115 (facade; evaluate(..., use_cuda))
  SystematicResampler  include/tsdf_localization/resampling/novel_resampling.h:38-74

Particles are ``float32[N, 7]`` arrays in the reference's ``Particle`` layout (x y z roll pitch yaw weight), points
``float32[P, 3]`` (``CudaPoint``). Everything numeric happens in libtsdfloc.so on the GPU; nothing here computes
weights on the CPU, and every call raises if the library or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import capi


def _f32(a, shape_tail: int, name: str) -> np.ndarray:
    arr = np.ascontiguousarray(a, dtype=np.float32)
    if arr.ndim != 2 or arr.shape[1] != shape_tail:
        raise ValueError(f"{name} must have shape [n, {shape_tail}]")
    return arr


def likelihood_value(tsdf_mm: float, sigma: float = 0.1) -> float:
    """TSDF value in mm -> N(d; 0, sigma)^3, createTSDFMap's transform (map_util.h:124-126)."""
    return float(capi.load_library().tsdfloc_likelihood_value(C.c_float(tsdf_mm), C.c_float(sigma)))


def likelihood_init(sigma: float = 0.1) -> float:
    """Value of unmapped space, N(10 m; 0, sigma)^3 (map_util.h:70-71)."""
    return float(capi.load_library().tsdfloc_likelihood_init(C.c_float(sigma)))


@dataclass
class PoseWithCovariance:
    """What the reference returns as geometry_msgs::PoseWithCovariance (cuda_evaluator.cu:410-423)."""
    position: tuple = (0.0, 0.0, 0.0)
    orientation: tuple = (0.0, 0.0, 0.0, 0.0)          # x y z w, from setRPY(roll, pitch, yaw)
    covariance: list = field(default_factory=lambda: [0.0] * 36)
    rpy: tuple = (0.0, 0.0, 0.0)

    @staticmethod
    def from_mean(mean6: Sequence[float]) -> "PoseWithCovariance":
        x, y, z, roll, pitch, yaw = (float(v) for v in mean6)
        hr, hp, hy = roll * 0.5, pitch * 0.5, yaw * 0.5
        cr, sr, cp, sp, cy, sy = math.cos(hr), math.sin(hr), math.cos(hp), math.sin(hp), math.cos(hy), math.sin(hy)
        q = (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)
        return PoseWithCovariance((x, y, z), q, [0.0] * 36, (roll, pitch, yaw))


class CudaSubVoxelMap:
    """Two-level sparse voxel map: 1 m upper cells -> dense sub_dim^3 fp32 bricks (host arrays)."""

    def __init__(self, min_x, min_y, min_z, max_x, max_y, max_z, resolution, init_value=0.0):
        self._lib = capi.load_library()
        self._h = C.c_void_p()
        mn = (C.c_float * 3)(min_x, min_y, min_z)
        mx = (C.c_float * 3)(max_x, max_y, max_z)
        rc = self._lib.tsdfloc_map_create(mn, mx, C.c_float(resolution), C.c_float(init_value), C.byref(self._h))
        if rc != capi.OK:
            raise ValueError("invalid map geometry")
        self._adopted = None

    @classmethod
    def from_arrays(cls, desc: capi.MapDesc, grid_occ: np.ndarray, data: np.ndarray) -> "CudaSubVoxelMap":
        """Adopt coef()/rawGridOcc()/rawData() of an existing (e.g. reference-built) map."""
        self = cls.__new__(cls)
        self._lib = capi.load_library()
        self._h = None
        d = capi.MapDesc()
        C.memmove(C.byref(d), C.byref(desc), C.sizeof(capi.MapDesc))
        self._adopted = (d, np.ascontiguousarray(grid_occ, dtype=np.int32), np.ascontiguousarray(data, dtype=np.float32))
        return self

    @classmethod
    def from_chunks(cls, chunk_pos, chunk_data, sigma: float = 0.1, device: Optional[int] = None) -> "CudaSubVoxelMap":
        """createTSDFMap (map_util.h:17-154) on chunks already read from the map file: chunk_pos int32[n, 3] (the integers of
        the dataset names "<cx>_<cy>_<cz>"), chunk_data uint32[n, 64, 64, 64] raw TSDFValue words. The free-space points are
        available as ``free_map()``. device=None: host ingest; device=k: the same ingest on GPU k (bit-identical arrays)."""
        pos = np.ascontiguousarray(chunk_pos, dtype=np.int32).reshape(-1, 3)
        dat = np.ascontiguousarray(chunk_data, dtype=np.uint32).reshape(len(pos), -1)
        if dat.shape[1] != 64 ** 3:
            raise ValueError("every chunk must hold 64^3 words")
        self = cls.__new__(cls)
        self._lib = capi.load_library()
        self._adopted = None
        self._h = C.c_void_p()
        if device is None:
            rc = self._lib.tsdfloc_map_from_chunks(pos.ctypes.data_as(C.c_void_p), dat.ctypes.data_as(C.c_void_p), len(pos), C.c_float(sigma),
                                                   C.byref(self._h))
        else:
            rc = self._lib.tsdfloc_map_from_chunks_gpu(pos.ctypes.data_as(C.c_void_p), dat.ctypes.data_as(C.c_void_p), len(pos),
                                                       C.c_float(sigma), int(device), C.byref(self._h))
        if rc == capi.E_CUDA:
            raise RuntimeError(self._lib.tsdfloc_last_error(None).decode())
        if rc != capi.OK:
            raise ValueError("invalid chunk set")
        return self

    def free_map(self) -> np.ndarray:
        if self._h is None:
            return np.zeros((0, 3), dtype=np.float32)
        n = C.c_uint64(0)
        p = self._lib.tsdfloc_map_free_points(self._h, C.byref(n))
        if not p or n.value == 0:
            return np.zeros((0, 3), dtype=np.float32)
        return np.ctypeslib.as_array(p, shape=(int(n.value), 3)).copy()

    def setData(self, cells) -> None:
        """cells: [n, 4] (x, y, z, value) — the tuple list createTSDFMap hands to setData (map_util.h:129,152)."""
        if self._h is None:
            raise RuntimeError("map adopted from arrays is read-only")
        arr = _f32(cells, 4, "cells")
        rc = self._lib.tsdfloc_map_set_data(self._h, arr.ctypes.data_as(C.POINTER(C.c_float)), arr.shape[0])
        if rc != capi.OK:
            raise ValueError("cell outside the map bounds (or map too large)")

    def coef(self) -> capi.MapDesc:
        if self._adopted is not None:
            return self._adopted[0]
        return self._lib.tsdfloc_map_get_desc(self._h).contents

    def rawGridOcc(self) -> np.ndarray:
        if self._adopted is not None:
            return self._adopted[1]
        n = int(self.coef().grid_occ_size)
        return np.ctypeslib.as_array(self._lib.tsdfloc_map_grid_occ(self._h), shape=(n,))

    def rawData(self) -> np.ndarray:
        if self._adopted is not None:
            return self._adopted[2]
        n = int(self.coef().data_size)
        if n == 0:
            return np.zeros(0, dtype=np.float32)
        return np.ctypeslib.as_array(self._lib.tsdfloc_map_data(self._h), shape=(n,))

    def dataBytes(self) -> int:
        return int(self.coef().data_size) * 4

    def gridOccBytes(self) -> int:
        return int(self.coef().grid_occ_size) * 4

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.tsdfloc_map_destroy(h)
            self._h = None


class CudaEvaluator:
    """GPU sensor update. Constructor uploads the map once (cuda_evaluator.cu:21-59)."""

    def __init__(self, map: CudaSubVoxelMap, per_point: bool = False, a_hit: float = 0.9, a_range: float = 0.1,
                 a_max: float = 0.0, max_range: float = 100.0, device: int = 0, neg_policy: int = capi.NEG_MISS):
        """neg_policy (not a reference argument): capi.NEG_MISS (default) or capi.NEG_SATURATE_LIKE_REF_GPU, which
        reproduces the reference CUDA build's handling of lookups below map.min bit for bit (tsdfloc.h)."""
        self._lib = capi.load_library()
        self._ctx = C.c_void_p()
        prm = capi.Params(a_hit, a_range, a_max, max_range, int(per_point), int(neg_policy))
        desc = map.coef()
        occ = np.ascontiguousarray(map.rawGridOcc(), dtype=np.int32)
        data = np.ascontiguousarray(map.rawData(), dtype=np.float32)
        rc = self._lib.tsdfloc_create(C.byref(desc), occ.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p),
                                      C.byref(prm), int(device), C.byref(self._ctx))
        if rc != capi.OK:
            msg = self._lib.tsdfloc_last_error(None).decode()
            # the reference wraps every constructor failure in this text (cuda_evaluator.cu:52-55)
            raise RuntimeError("Error while creating the CUDA context for the map! " + msg)
        self.device = device
        self.data_size = int(desc.data_size)

    @classmethod
    def from_chunks(cls, chunk_pos, chunk_data, sigma: float = 0.1, per_point: bool = False, a_hit: float = 0.9, a_range: float = 0.1,
                    a_max: float = 0.0, max_range: float = 100.0, device: int = 0, neg_policy: int = capi.NEG_MISS) -> "CudaEvaluator":
        """createTSDFMap + the constructor in one step on the GPU (tsdfloc_create_from_chunks): the bricks are built on
        `device` and never visit the host; the free-space points stay there for ParticleCloud.initialize(free-map mode)."""
        pos = np.ascontiguousarray(chunk_pos, dtype=np.int32).reshape(-1, 3)
        dat = np.ascontiguousarray(chunk_data, dtype=np.uint32).reshape(len(pos), -1)
        if dat.shape[1] != 64 ** 3:
            raise ValueError("every chunk must hold 64^3 words")
        self = cls.__new__(cls)
        self._lib = capi.load_library()
        self._ctx = C.c_void_p()
        prm = capi.Params(a_hit, a_range, a_max, max_range, int(per_point), int(neg_policy))
        rc = self._lib.tsdfloc_create_from_chunks(pos.ctypes.data_as(C.c_void_p), dat.ctypes.data_as(C.c_void_p), len(pos), C.c_float(sigma),
                                                  C.byref(prm), int(device), C.byref(self._ctx))
        if rc != capi.OK:
            raise RuntimeError("Error while creating the CUDA context for the map! " + self._lib.tsdfloc_last_error(None).decode())
        self.device = device
        self.data_size = int(self.map_desc().data_size)
        return self

    def map_desc(self) -> capi.MapDesc:
        d = capi.MapDesc()
        capi.check(self._lib, self._ctx, self._lib.tsdfloc_map_desc_of(self._ctx, C.byref(d)))
        return d

    def free_map_size(self) -> int:
        """Free-space points resident on the device (maps ingested by from_chunks)."""
        p, n = C.c_void_p(), C.c_uint64(0)
        capi.check(self._lib, self._ctx, self._lib.tsdfloc_free_map_device(self._ctx, C.byref(p), C.byref(n)))
        return int(n.value)

    # -- reference interface -------------------------------------------------------------------------------------
    def evaluate(self, particles: np.ndarray, points, tf_matrix) -> PoseWithCovariance:
        """evaluate(std::vector<Particle>&, const std::vector<CudaPoint>&, FLOAT_T tf_matrix[16]).

        Writes the NORMALISED weights into ``particles[:, 6]`` in place and returns the weighted mean pose.
        An empty scan returns a default pose and leaves the weights untouched (cuda_evaluator.cu:122-125).
        Raises RuntimeError("No particle is valid!") when all weights are zero (cuda_evaluator.cu:366-369).
        """
        if not (isinstance(particles, np.ndarray) and particles.dtype == np.float32 and particles.ndim == 2
                and particles.shape[1] == 7 and particles.flags.c_contiguous):
            raise ValueError("particles must be a C-contiguous float32[n, 7] array (it is updated in place)")
        pts = _f32(np.asarray(points).reshape(-1, 3), 3, "points")
        if pts.shape[0] == 0:
            return PoseWithCovariance()
        tf = (C.c_float * 16)(*[float(v) for v in np.asarray(tf_matrix, dtype=np.float32).reshape(-1)[:16]])
        mean = (C.c_float * 6)()
        rc = self._lib.tsdfloc_sensor_update(self._ctx, particles.ctypes.data_as(C.c_void_p), particles.shape[0],
                                             pts.ctypes.data_as(C.c_void_p), pts.shape[0], tf, mean)
        if rc == capi.E_NO_VALID_PARTICLE:
            raise RuntimeError("No particle is valid!")
        capi.check(self._lib, self._ctx, rc)
        return PoseWithCovariance.from_mean(list(mean))

    # -- scan reduction + evaluation on a raw cloud (TSDFEvaluator::evaluateParticles, tsdf_evaluator.cpp:304-378) ----------
    @staticmethod
    def _cloud_args(points, ring):
        pts = _f32(np.asarray(points).reshape(-1, 3), 3, "points")
        if ring is None:
            return pts, None, (pts.ctypes.data_as(C.c_void_p), 12, None, 0, 4, pts.shape[0])
        rg = np.ascontiguousarray(ring)
        if rg.dtype not in (np.int16, np.int32):
            rg = rg.astype(np.int32)
        if rg.shape != (pts.shape[0],):
            raise ValueError("ring must have one entry per point")
        return pts, rg, (pts.ctypes.data_as(C.c_void_p), 12, rg.ctypes.data_as(C.c_void_p), rg.itemsize, rg.itemsize, pts.shape[0])

    def reduce_scan(self, points, ring, cell_size: float = 0.064, n_rings: int = 128, ring_desync_like_reference: bool = False,
                    emit_centres: bool = False, want_src: bool = False):
        """The reference's scan reduction on the GPU: drop points nearer than 1 m, keep the first point per (ring, cell),
        emit the original points ring-major in cloud order (tsdf_evaluator.cpp:304-376). ring=None: one ring."""
        pts, rg, args = self._cloud_args(points, ring)
        n = pts.shape[0]
        flags = (capi.REDUCE_RING_DESYNC_LIKE_REFERENCE if ring_desync_like_reference else 0) | (capi.REDUCE_EMIT_CENTRES if emit_centres else 0)
        out = np.empty((max(n, 1), 3), dtype=np.float32)
        src = np.empty(max(n, 1), dtype=np.uint32) if want_src else None
        n_out = C.c_uint64(0)
        rc = self._lib.tsdfloc_reduce_scan(self._ctx, *args, C.c_double(cell_size), n_rings, flags, out.ctypes.data_as(C.c_void_p),
                                           src.ctypes.data_as(C.c_void_p) if want_src else None, n, C.byref(n_out))
        capi.check(self._lib, self._ctx, rc)
        m = int(n_out.value)
        return (out[:m], src[:m]) if want_src else out[:m]

    def evaluate_cloud(self, particles: np.ndarray, points, ring, tf_matrix, cell_size: float = 0.064, n_rings: int = 128,
                       ring_desync_like_reference: bool = False):
        """Reduction + evaluate() in one call; the reduced scan stays on the device. Returns (pose, reduced scan size)."""
        if not (isinstance(particles, np.ndarray) and particles.dtype == np.float32 and particles.ndim == 2
                and particles.shape[1] == 7 and particles.flags.c_contiguous):
            raise ValueError("particles must be a C-contiguous float32[n, 7] array (it is updated in place)")
        pts, rg, args = self._cloud_args(points, ring)
        tf = (C.c_float * 16)(*[float(v) for v in np.asarray(tf_matrix, dtype=np.float32).reshape(-1)[:16]])
        mean = (C.c_float * 6)()
        used = C.c_uint64(0)
        flags = capi.REDUCE_RING_DESYNC_LIKE_REFERENCE if ring_desync_like_reference else 0
        rc = self._lib.tsdfloc_sensor_update_cloud(self._ctx, particles.ctypes.data_as(C.c_void_p), particles.shape[0], *args,
                                                   C.c_double(cell_size), n_rings, flags, tf, mean, C.byref(used))
        if rc == capi.E_EMPTY_SCAN:
            return PoseWithCovariance(), 0
        if rc == capi.E_NO_VALID_PARTICLE:
            raise RuntimeError("No particle is valid!")
        capi.check(self._lib, self._ctx, rc)
        return PoseWithCovariance.from_mean(list(mean)), int(used.value)

    def best_particle(self):
        """(index, pose6, weight) of the first particle carrying the largest weight of the last evaluate()
        (src/mcl_3d.cpp:382-399); index -1 when no weight is > 0."""
        idx = C.c_int64(-1)
        pose = (C.c_float * 6)()
        w = C.c_float(0.0)
        capi.check(self._lib, self._ctx, self._lib.tsdfloc_best_particle(self._ctx, C.byref(idx), pose, C.byref(w), None))
        return int(idx.value), np.array(list(pose), dtype=np.float32), float(w.value)

    # -- resampling on the particle set the last evaluate() left on the device -------------------------------------
    def resample_systematic(self, u0: float, capacity: Optional[int] = None, want_parents: bool = False):
        n = capacity if capacity is not None else 0
        if n <= 0:
            raise ValueError("capacity must be positive")
        out = np.empty((n, 7), dtype=np.float32)
        parents = np.empty(n, dtype=np.uint32) if want_parents else None
        n_out = C.c_uint64(0)
        rc = self._lib.tsdfloc_resample_systematic(self._ctx, C.c_float(u0), out.ctypes.data_as(C.c_void_p), n, C.byref(n_out),
                                                   parents.ctypes.data_as(C.c_void_p) if want_parents else None)
        capi.check(self._lib, self._ctx, rc)
        m = int(n_out.value)
        return (out[:m], parents[:m]) if want_parents else out[:m]

    # -- parity/debug -------------------------------------------------------------------------------------------------
    def debug_eval(self, particles, points, tf_matrix, want_idx: bool = True):
        """Per-pair flat voxel indices (data_size = miss), per-particle hit counts and un-normalised weights."""
        ps = _f32(particles, 7, "particles")
        pts = _f32(points, 3, "points")
        n, p = ps.shape[0], pts.shape[0]
        tf = (C.c_float * 16)(*[float(v) for v in np.asarray(tf_matrix, dtype=np.float32).reshape(-1)[:16]])
        idx = np.empty((n, p), dtype=np.uint32) if want_idx else None
        hits = np.empty(n, dtype=np.uint32)
        raw = np.empty(n, dtype=np.float32)
        rc = self._lib.tsdfloc_debug_eval(self._ctx, ps.ctypes.data_as(C.c_void_p), n, pts.ctypes.data_as(C.c_void_p), p, tf,
                                          idx.ctypes.data_as(C.c_void_p) if want_idx else None,
                                          hits.ctypes.data_as(C.c_void_p), raw.ctypes.data_as(C.c_void_p))
        capi.check(self._lib, self._ctx, rc)
        return idx, hits, raw

    @property
    def ctx(self) -> C.c_void_p:
        return self._ctx

    def kernel_launches(self) -> int:
        return int(self._lib.tsdfloc_kernel_launches(self._ctx))

    def tune(self, knob: int, value: int) -> None:
        """Test / tuning hook (tsdfloc_tune): every setting produces the same bits."""
        capi.check(self._lib, self._ctx, self._lib.tsdfloc_tune(self._ctx, int(knob), int(value)))

    def graph_stats(self):
        """(recordings made, updates served by a graph launch, why the last recording attempt was abandoned)."""
        st = (C.c_uint64 * 2)()
        note = C.c_char_p()
        capi.check(self._lib, self._ctx, self._lib.tsdfloc_graph_stats(self._ctx, st, C.byref(note)))
        return int(st[0]), int(st[1]), (note.value or b"").decode()

    def division_mode(self):
        """(mode, open brackets): what tsdfloc_create proved for the map's resolution (capi.DIV_*)."""
        n = C.c_uint64(0)
        return int(self._lib.tsdfloc_division_mode(self._ctx, C.byref(n))), int(n.value)

    def eval_stats(self):
        st = (C.c_uint64 * 4)()
        capi.check(self._lib, self._ctx, self._lib.tsdfloc_eval_stats(self._ctx, st))
        return dict(blocks=int(st[0]), folded=int(st[1]), tie_folds=int(st[2]), redone=int(st[3]))

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._lib.tsdfloc_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        self.close()


class MultiGpuEvaluator:
    """CudaEvaluator over several GPUs of one box from ONE process (tsdfloc_multi_*): same evaluate() / resample_systematic()
    contract and bit-identical results; particles are sharded, map and scan replicated, and the kernels store their results
    straight into every device's buffers over NVLink. ``devices`` may repeat a device (tests on a one-GPU machine)."""

    def __init__(self, map: CudaSubVoxelMap, devices: Sequence[int], per_point: bool = False, a_hit: float = 0.9, a_range: float = 0.1,
                 a_max: float = 0.0, max_range: float = 100.0):
        self._lib = capi.load_library()
        self._m = C.c_void_p()
        prm = capi.Params(a_hit, a_range, a_max, max_range, int(per_point), 0)
        desc = map.coef()
        occ = np.ascontiguousarray(map.rawGridOcc(), dtype=np.int32)
        data = np.ascontiguousarray(map.rawData(), dtype=np.float32)
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        rc = self._lib.tsdfloc_multi_create(C.byref(desc), occ.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p), C.byref(prm),
                                            devs, len(devices), C.byref(self._m))
        if rc != capi.OK:
            raise RuntimeError("Error while creating the CUDA context for the map! " + self._lib.tsdfloc_multi_last_error(None).decode())
        self.devices = list(devices)

    def _check(self, rc: int) -> None:
        if rc == capi.OK:
            return
        if rc == capi.E_NO_VALID_PARTICLE:
            raise RuntimeError("No particle is valid!")
        raise capi.TsdflocError(rc, self._lib.tsdfloc_multi_last_error(self._m).decode())

    def evaluate(self, particles: np.ndarray, points, tf_matrix) -> PoseWithCovariance:
        if not (isinstance(particles, np.ndarray) and particles.dtype == np.float32 and particles.ndim == 2
                and particles.shape[1] == 7 and particles.flags.c_contiguous):
            raise ValueError("particles must be a C-contiguous float32[n, 7] array (it is updated in place)")
        pts = _f32(np.asarray(points).reshape(-1, 3), 3, "points")
        if pts.shape[0] == 0:
            return PoseWithCovariance()
        tf = (C.c_float * 16)(*[float(v) for v in np.asarray(tf_matrix, dtype=np.float32).reshape(-1)[:16]])
        mean = (C.c_float * 6)()
        self._check(self._lib.tsdfloc_multi_sensor_update(self._m, particles.ctypes.data_as(C.c_void_p), particles.shape[0],
                                                          pts.ctypes.data_as(C.c_void_p), pts.shape[0], tf, mean))
        return PoseWithCovariance.from_mean(list(mean))

    def evaluate_cloud(self, particles: np.ndarray, points, ring, tf_matrix, cell_size: float = 0.064, n_rings: int = 128,
                       ring_desync_like_reference: bool = False):
        """TSDFEvaluator::evaluateParticles over all devices: the first device reduces the raw cloud, the reduced scan goes to the
        others over NVLink, then the sharded update. Returns (pose, reduced scan size) like CudaEvaluator.evaluate_cloud."""
        if not (isinstance(particles, np.ndarray) and particles.dtype == np.float32 and particles.ndim == 2
                and particles.shape[1] == 7 and particles.flags.c_contiguous):
            raise ValueError("particles must be a C-contiguous float32[n, 7] array (it is updated in place)")
        pts, rg, args = CudaEvaluator._cloud_args(points, ring)
        tf = (C.c_float * 16)(*[float(v) for v in np.asarray(tf_matrix, dtype=np.float32).reshape(-1)[:16]])
        mean = (C.c_float * 6)()
        used = C.c_uint64(0)
        flags = capi.REDUCE_RING_DESYNC_LIKE_REFERENCE if ring_desync_like_reference else 0
        rc = self._lib.tsdfloc_multi_sensor_update_cloud(self._m, particles.ctypes.data_as(C.c_void_p), particles.shape[0], *args,
                                                         C.c_double(cell_size), n_rings, flags, tf, mean, C.byref(used))
        if rc == capi.E_EMPTY_SCAN:
            return PoseWithCovariance(), 0
        self._check(rc)
        return PoseWithCovariance.from_mean(list(mean)), int(used.value)

    def resample_systematic(self, u0: float, capacity: int):
        out = np.empty((capacity, 7), dtype=np.float32)
        n_out = C.c_uint64(0)
        self._check(self._lib.tsdfloc_multi_resample_systematic(self._m, C.c_float(u0), out.ctypes.data_as(C.c_void_p), capacity, C.byref(n_out)))
        return out[:int(n_out.value)]

    def close(self) -> None:
        if getattr(self, "_m", None):
            self._lib.tsdfloc_multi_destroy(self._m)
            self._m = None

    def __del__(self):
        self.close()


class ParticleCloud:
    """Motion-update half of the reference's ParticleCloud (include/tsdf_localization/particle_cloud.h:44-258,
    src/particle_cloud.cpp:153-617) on the GPU: the four ``motionUpdate`` variants, the reference pose that gates the sensor
    update (``refDist`` / ``refAngle`` / ``resetRef``) and the motion parameters a_1..a_12. ``particles`` is a float32[n, 7]
    array updated in place. The reference measures ``time_diff`` with ros::Time::now(); here the caller passes it.

    Samples come from a counter-based Philox stream on the device (``seed``; one sequence number per update), or — parity
    mode — from ``draws`` (float64[n, 6]), the values the reference's six std::normal_distribution<> objects returned.
    """

    def __init__(self, evaluator, particles: Optional[np.ndarray] = None, seed: int = 0):
        self._ev = evaluator.cuda_evaluator_ if isinstance(evaluator, TSDFEvaluator) else evaluator
        self._lib = capi.load_library()
        self.m_particles = particles
        self.ref_pose = np.zeros(6, dtype=np.float32)
        self.a = np.full(12, 0.1, dtype=np.float32)           # particle_cloud.h:57-68
        self.seed = int(seed)
        self.sequence = 0

    def particles(self) -> np.ndarray:
        return self.m_particles

    def size(self) -> int:
        return 0 if self.m_particles is None else len(self.m_particles)

    def setAParams(self, *a) -> None:
        if len(a) != 12:
            raise ValueError("a_1 .. a_12 expected")
        self.a = np.asarray(a, dtype=np.float32)

    def refDist(self) -> float:
        r = self.ref_pose
        return float(np.sqrt(np.float32(r[0] * r[0] + r[1] * r[1] + r[2] * r[2])))

    def refAngle(self) -> float:
        return float(abs(self.ref_pose[5]))

    def resetRef(self) -> None:
        self.ref_pose[:] = 0

    def initialize(self, number_particles: int, center_pose, spread, mode: int = capi.INIT_NORMAL, free_map=None) -> np.ndarray:
        """The reference's three initialize() overloads (particle_cloud.cpp:32-148): center_pose = x y z roll pitch yaw,
        spread = sigmas (INIT_NORMAL) or half-widths (INIT_UNIFORM / INIT_FREE_MAP; the free-map mode ignores the xyz entries).
        Resets the reference pose, returns (and keeps) the new float32[n, 7] particle array."""
        self.ref_pose[:] = 0
        ps = np.empty((int(number_particles), 7), dtype=np.float32)
        mean = (C.c_double * 6)(*[float(v) for v in center_pose])
        spr = (C.c_double * 6)(*[float(v) for v in spread])
        fm, n_free = None, 0
        if free_map is not None:
            fm = _f32(free_map, 3, "free_map")
            n_free = fm.shape[0]
        self.sequence += 1
        capi.check(self._lib, self._ev.ctx, self._lib.tsdfloc_init_particles(
            self._ev.ctx, ps.ctypes.data_as(C.c_void_p), len(ps), int(mode), mean, spr, fm.ctypes.data_as(C.c_void_p) if fm is not None else None,
            n_free, self.seed, self.sequence))
        self.m_particles = ps
        return ps

    def model(self, variant: int, inputs, time_diff: float):
        """(mean[6], sigma[6]) of the variant's six normal distributions; advances ref_pose like the reference."""
        inp = (C.c_double * 4)(*([float(v) for v in inputs] + [0.0] * (4 - len(inputs))))
        mean, sigma = (C.c_double * 6)(), (C.c_double * 6)()
        rc = self._lib.tsdfloc_motion_model(int(variant), inp, C.c_float(time_diff), self.a.ctypes.data_as(C.POINTER(C.c_float)), mean,
                                            sigma, self.ref_pose.ctypes.data_as(C.POINTER(C.c_float)))
        if rc != capi.OK:
            raise ValueError("unknown motion variant")
        return np.array(list(mean)), np.array(list(sigma))

    def _apply(self, mean, sigma, draws):
        ps = self.m_particles
        if not (isinstance(ps, np.ndarray) and ps.dtype == np.float32 and ps.ndim == 2 and ps.shape[1] == 7 and ps.flags.c_contiguous):
            raise ValueError("particles must be a C-contiguous float32[n, 7] array (it is updated in place)")
        m = (C.c_double * 6)(*mean)
        s = (C.c_double * 6)(*sigma)
        dptr = None
        if draws is not None:
            draws = np.ascontiguousarray(draws, dtype=np.float64)
            if draws.shape != (len(ps), 6):
                raise ValueError("draws must have shape [n, 6]")
            dptr = draws.ctypes.data_as(C.c_void_p)
        self.sequence += 1
        capi.check(self._lib, self._ev.ctx, self._lib.tsdfloc_motion_update(self._ev.ctx, ps.ctypes.data_as(C.c_void_p), len(ps), m, s, dptr,
                                                                           self.seed, self.sequence))

    # the reference's four motionUpdate overloads (particle_cloud.cpp:153, 333, 388, 422), one method each
    def motionUpdateNoise(self, lin_scale: float, ang_scale: float, time_diff: float, draws=None) -> None:
        self._apply(*self.model(capi.MOTION_NOISE, [lin_scale, ang_scale], time_diff), draws)

    def motionUpdateOdom(self, linear_x: float, angular_z: float, time_diff: float, draws=None) -> None:
        self._apply(*self.model(capi.MOTION_ODOM, [linear_x, angular_z], time_diff), draws)

    def motionUpdateImu(self, linear_vel: float, angular_yaw: float, time_diff: float, draws=None) -> None:
        self._apply(*self.model(capi.MOTION_IMU, [linear_vel, angular_yaw], time_diff), draws)

    def motionUpdateNoiseImu(self, lin_scale: float, delta_roll: float, delta_pitch: float, delta_yaw: float, time_diff: float,
                             draws=None) -> None:
        self._apply(*self.model(capi.MOTION_NOISE_IMU, [lin_scale, delta_roll, delta_pitch, delta_yaw], time_diff), draws)


class TSDFEvaluator:
    """Sensor-update facade with the reference's constructor and evaluate() signature (tsdf_evaluator.h:72-114).

    The reference dispatches ``use_cuda`` between its CUDA evaluator and a CPU/OpenMP loop; this build ships only the
    GPU path, so ``use_cuda=False`` raises instead of silently computing on the host.
    """

    def __init__(self, map_ptr: CudaSubVoxelMap, per_point: bool = False, a_hit: float = 0.9, a_range: float = 0.1,
                 a_max: float = 0.0, max_range: float = 100.0, reduction_cell_size: float = 0.064, device: int = 0):
        self.map_ptr_ = map_ptr
        self.cuda_evaluator_ = CudaEvaluator(map_ptr, per_point, a_hit, a_range, a_max, max_range, device=device)
        self.map_res_ = reduction_cell_size

    def evaluate(self, particles: np.ndarray, points, tf_matrix, use_cuda: bool = True) -> PoseWithCovariance:
        if not use_cuda:
            raise RuntimeError("tsdf_localization_b200 has no CPU evaluator: call evaluate(..., use_cuda=True)")
        return self.cuda_evaluator_.evaluate(particles, points, tf_matrix)

    def evaluateParticles(self, particles: np.ndarray, points, ring, tf_matrix=None, use_cuda: bool = True, n_rings: int = 128,
                          ring_desync_like_reference: bool = False) -> PoseWithCovariance:
        """evaluateParticles(particle_cloud, real_cloud, ..., use_cuda, ignore_tf) (tsdf_evaluator.cpp:247-378) with the cloud
        given as points + per-point ring and the scanner->robot transform as a matrix (None = ignore_tf): GPU scan reduction
        at reduction_cell_size, then evaluate()."""
        if not use_cuda:
            raise RuntimeError("tsdf_localization_b200 has no CPU evaluator: call evaluateParticles(..., use_cuda=True)")
        tf = np.eye(4, dtype=np.float32).reshape(-1) if tf_matrix is None else tf_matrix
        pose, _ = self.cuda_evaluator_.evaluate_cloud(particles, points, ring, tf, self.map_res_, n_rings, ring_desync_like_reference)
        return pose


class SystematicResampler:
    """Systematic resampling on the GPU with the reference's recurrence (novel_resampling.h:41-72).

    ``resample(particle_cloud)`` mirrors ``Resampler::resample(ParticleCloud&)`` (resampler.h:26): a weighted particle set
    in, the resampled set out (its length is whatever the reference recurrence emits — usually n);
    copies keep the parent's weight. The reference draws U0 ~ uniform_real_distribution<float>(0, 1/N) from its own
    std::mt19937; here U0 comes from a seeded numpy Generator or is passed explicitly (parity is defined "given the
    same weights and the same U0"). The GPU context (device, stream, scratch) is borrowed from an evaluator.
    """

    def __init__(self, evaluator, seed: Optional[int] = None):
        self._ev = evaluator.cuda_evaluator_ if isinstance(evaluator, TSDFEvaluator) else evaluator
        self._rng = np.random.default_rng(seed)

    def draw_u0(self, n: int) -> float:
        inv_m = 1.0 / n
        u = np.float32(self._rng.random() * inv_m)
        if float(u) >= inv_m:   # rounding up to the open bound
            u = np.nextafter(u, np.float32(0.0))
        return float(u)

    def resample(self, particle_cloud: np.ndarray, u0: Optional[float] = None, want_parents: bool = False):
        ps = _f32(particle_cloud, 7, "particle_cloud")
        n = ps.shape[0]
        if u0 is None:
            u0 = self.draw_u0(n)
        cap = n + n // 8 + 64
        out = np.empty((cap, 7), dtype=np.float32)
        parents = np.empty(cap, dtype=np.uint32) if want_parents else None
        n_out = C.c_uint64(0)
        lib = self._ev._lib
        rc = lib.tsdfloc_resample_particles(self._ev.ctx, ps.ctypes.data_as(C.c_void_p), n, C.c_float(u0),
                                            out.ctypes.data_as(C.c_void_p), cap, C.byref(n_out),
                                            parents.ctypes.data_as(C.c_void_p) if want_parents else None)
        capi.check(lib, self._ev.ctx, rc)
        m = int(n_out.value)
        return (out[:m], parents[:m]) if want_parents else out[:m]

    def resample_resident(self, n: int, u0: Optional[float] = None, want_parents: bool = False):
        """Resample the particle set the evaluator's last evaluate() left on the device (no re-upload)."""
        if u0 is None:
            u0 = self.draw_u0(n)
        return self._ev.resample_systematic(u0, capacity=n + n // 8 + 64, want_parents=want_parents)


class _RunResampler:
    """Shared plumbing of the two resamplers whose recurrence runs on the host (tsdfloc_resample)."""

    METHOD = None

    def __init__(self, evaluator, seed: Optional[int] = None):
        self._ev = evaluator.cuda_evaluator_ if isinstance(evaluator, TSDFEvaluator) else evaluator
        self._rng = np.random.default_rng(seed)

    def _call(self, particle_cloud, n, u, draw_cb, cap, want_parents):
        lib = self._ev._lib
        ps = None if particle_cloud is None else _f32(particle_cloud, 7, "particle_cloud")
        if ps is not None:
            n = ps.shape[0]
        out = np.empty((cap, 7), dtype=np.float32)
        parents = np.empty(cap, dtype=np.uint32) if want_parents else None
        n_out = C.c_uint64(0)
        rc = lib.tsdfloc_resample(self._ev.ctx, self.METHOD, ps.ctypes.data_as(C.c_void_p) if ps is not None else None, n, C.c_float(u),
                                  draw_cb if draw_cb is not None else C.cast(None, capi.INDEX_DRAW_FN), None,
                                  out.ctypes.data_as(C.c_void_p), cap, C.byref(n_out),
                                  parents.ctypes.data_as(C.c_void_p) if want_parents else None)
        capi.check(lib, self._ev.ctx, rc)
        m = int(n_out.value)
        return (out[:m], parents[:m]) if want_parents else out[:m]


class ResidualSystematicResampler(_RunResampler):
    """ResidualSystematicResampler (novel_resampling.h:76-104), mcl_3d's compiled-in default (src/mcl_3d.cpp:765): the fp32
    remainder recurrence runs on the host over the weights, the copies are made on the GPU. ``u0`` is the reference's
    uniform_real_distribution<float>(0, 1) draw (from a seeded numpy Generator when not given)."""

    METHOD = capi.RESAMPLE_RESIDUAL_SYSTEMATIC

    def resample(self, particle_cloud: np.ndarray, u0: Optional[float] = None, want_parents: bool = False):
        u = float(np.float32(self._rng.random())) if u0 is None else u0
        return self._call(particle_cloud, 0, u, None, 2 * len(particle_cloud) + 64, want_parents)

    def resample_resident(self, n: int, u0: Optional[float] = None, want_parents: bool = False):
        u = float(np.float32(self._rng.random())) if u0 is None else u0
        return self._call(None, n, u, None, 2 * n + 64, want_parents)


class ResidualResampler(_RunResampler):
    """ResidualResampler (novel_resampling.h:9-36), the dynamic-reconfigure default (cfg/MCL.cfg:54): uniformly drawn
    particles are copied ceil(w * N) times until N slots are filled. ``index_draws``: the uniform index stream to consume (the
    reference's std::uniform_int_distribution<size_t>(0, N-1) draws, for parity); default: a seeded numpy Generator."""

    METHOD = capi.RESAMPLE_RESIDUAL

    def _draw_cb(self, n, index_draws):
        if index_draws is None:
            rng = self._rng
            return capi.INDEX_DRAW_FN(lambda _user: int(rng.integers(0, n)))
        it = iter(np.asarray(index_draws, dtype=np.uint64).tolist())
        return capi.INDEX_DRAW_FN(lambda _user: next(it, n))      # n = "out of draws": rejected by the library as a bad draw

    def resample(self, particle_cloud: np.ndarray, index_draws=None, want_parents: bool = False):
        n = len(particle_cloud)
        return self._call(particle_cloud, 0, 0.0, self._draw_cb(n, index_draws), n, want_parents)

    def resample_resident(self, n: int, index_draws=None, want_parents: bool = False):
        return self._call(None, n, 0.0, self._draw_cb(n, index_draws), n, want_parents)



class DrawSource:
    """The random draws a drawn resampler consumes, as the C callbacks of ``tsdfloc_draws``.

    ``real`` / ``index`` are either C function pointers taking ``user`` (a generator living in native code — how the C++ shim
    hands over the reference's std::mt19937) or Python callables (wrapped here; slow, for small sets). ``DrawSource.numpy(seed,
    n)`` draws from a seeded numpy Generator."""

    def __init__(self, real, index=None, user=None, wheel_real=None):
        self._keep = []
        self.real = self._wrap(real, capi.REAL_DRAW_FN)
        self.wheel_real = self._wrap(wheel_real, capi.REAL_DRAW_FN) if wheel_real is not None else self.real
        self.index = self._wrap(index, capi.INDEX_DRAW_FN) if index is not None else C.cast(None, capi.INDEX_DRAW_FN)
        self.user = user

    def _wrap(self, fn, proto):
        if isinstance(fn, proto):
            return fn
        if isinstance(fn, C._CFuncPtr):                      # a symbol of another shared library
            return C.cast(fn, proto)
        cb = proto(lambda _user: fn())
        self._keep.append(cb)
        return cb

    @classmethod
    def numpy(cls, seed, n):
        rng = np.random.default_rng(seed)
        return cls(real=lambda: float(np.float32(rng.random(dtype=np.float32))), index=lambda: int(rng.integers(0, n)))


class _DrawnResampler:
    """Shared plumbing of the resamplers that pick every output particle from random draws (tsdfloc_resample_drawn)."""

    METHOD = None
    STEPS = 0          # sampling_steps_ (Metropolis only)

    def __init__(self, evaluator, seed: Optional[int] = None):
        self._ev = evaluator.cuda_evaluator_ if isinstance(evaluator, TSDFEvaluator) else evaluator
        self._seed = seed

    def _call(self, particle_cloud, n, draws, steps, max_draws, want_parents):
        lib = self._ev._lib
        ps = None if particle_cloud is None else _f32(particle_cloud, 7, "particle_cloud")
        if ps is not None:
            n = ps.shape[0]
        if draws is None:
            draws = DrawSource.numpy(self._seed, n)
        d = capi.Draws(draws.wheel_real if self.METHOD == capi.RESAMPLE_WHEEL else draws.real, draws.index, draws.user, steps, max_draws)
        out = np.empty((n, 7), dtype=np.float32)
        parents = np.empty(n, dtype=np.uint32) if want_parents else None
        n_out = C.c_uint64(0)
        rc = lib.tsdfloc_resample_drawn(self._ev.ctx, self.METHOD, ps.ctypes.data_as(C.c_void_p) if ps is not None else None, n, C.byref(d),
                                        out.ctypes.data_as(C.c_void_p), n, C.byref(n_out),
                                        parents.ctypes.data_as(C.c_void_p) if want_parents else None)
        capi.check(lib, self._ev.ctx, rc)
        m = int(n_out.value)
        return (out[:m], parents[:m]) if want_parents else out[:m]

    def resample(self, particle_cloud: np.ndarray, draws: Optional[DrawSource] = None, want_parents: bool = False, max_draws: int = 0):
        return self._call(particle_cloud, 0, draws, self.STEPS, max_draws, want_parents)

    def resample_resident(self, n: int, draws: Optional[DrawSource] = None, want_parents: bool = False, max_draws: int = 0):
        """Resample the particle set the evaluator's last evaluate() left on the device (only the weights visit the host)."""
        return self._call(None, n, draws, self.STEPS, max_draws, want_parents)


class WheelResampler(_DrawnResampler):
    """WheelResampler (src/resampling/wheel_resampler.cpp:6-34), case 0 of mcl_3d's switch (src/mcl_3d.cpp:245-247): every slot
    draws u and takes the first particle whose fp32 running weight sum reaches it. O(n) here (one prefix + guide table) instead
    of the reference's O(n^2) walk; same parents given the same draws."""

    METHOD = capi.RESAMPLE_WHEEL


class MetropolisResampler(_DrawnResampler):
    """MetropolisResampler(sampling_steps) (novel_resampling.h:106-144), case 4 (the node passes 50, src/mcl_3d.cpp:257-259)."""

    METHOD = capi.RESAMPLE_METROPOLIS

    def __init__(self, evaluator, sampling_steps: int = 50, seed: Optional[int] = None):
        super().__init__(evaluator, seed)
        self.STEPS = int(sampling_steps)


class RejectionResampler(_DrawnResampler):
    """RejectionResampler (novel_resampling.h:146-189), the switch's default branch (src/mcl_3d.cpp:260-262). ``max_draws``
    bounds the index draws (0 = unbounded like the reference, which spins on e.g. all-negative weights)."""

    METHOD = capi.RESAMPLE_REJECTION
