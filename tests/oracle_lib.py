"""ctypes access to the CPU checkers under oracle/ (TEST INFRASTRUCTURE; never imported by the product package).

  Oracle  — oracle/_build/libtsdf_oracle.so, the plain-C restatement (always available, built by oracle/Makefile)
  Ref     — oracle/_ref/libtsdf_ref*.so, the UNMODIFIED reference CPU evaluator/map/resampler compiled against stub ROS
            headers (built in the container that has /root/reference; the prebuilt .so travels to the GPU box)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"

NEG_AS_MISS, NEG_REF_HOST_X86, NEG_REF_DEVICE_SAT = 0, 1, 2


class Coef(C.Structure):
    _fields_ = [
        ("dim", C.c_uint64 * 3), ("min", C.c_float * 3), ("max", C.c_float * 3), ("resolution", C.c_float),
        ("init_value", C.c_float), ("up_dim", C.c_uint64 * 3), ("up_dim_2", C.c_uint64), ("sub_dim", C.c_uint64),
        ("sub_dim_2", C.c_uint64), ("grid_occ_size", C.c_uint64), ("data_size", C.c_uint64),
    ]


class OracleMapStruct(C.Structure):
    _fields_ = [("coef", Coef), ("grid_occ", C.POINTER(C.c_int32)), ("data", C.POINTER(C.c_float))]


class OParams(C.Structure):
    _fields_ = [("a_hit", C.c_float), ("a_range", C.c_float), ("a_max", C.c_float), ("max_range", C.c_float)]


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


def build_oracle():
    so = ORACLE_DIR / "_build" / "libtsdf_oracle.so"
    src = [ORACLE_DIR / "tsdf_oracle.c", ORACLE_DIR / "tsdf_oracle.h"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in src):
        subprocess.run(["make", "-C", str(ORACLE_DIR), "oracle"], check=True, capture_output=True)
    return so


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(str(build_oracle()))
        L = self.lib
        L.oracle_map_create.restype = C.POINTER(OracleMapStruct)
        L.oracle_map_create.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float]
        L.oracle_map_from_arrays.restype = C.POINTER(OracleMapStruct)
        L.oracle_map_from_arrays.argtypes = [C.POINTER(Coef), C.c_void_p, C.c_void_p]
        L.oracle_map_destroy.argtypes = [C.c_void_p]
        L.oracle_map_set_data.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.oracle_get_index.restype = C.c_uint64
        L.oracle_get_index.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int]
        L.oracle_get_entries.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        L.oracle_likelihood_value.restype = C.c_float
        L.oracle_likelihood_value.argtypes = [C.c_float, C.c_float]
        L.oracle_likelihood_init.restype = C.c_float
        L.oracle_likelihood_init.argtypes = [C.c_float]
        L.oracle_pose_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_pose_weight.restype = C.c_float
        L.oracle_pose_weight.argtypes = [C.c_void_p, C.POINTER(OParams), C.c_void_p, C.c_void_p, C.c_uint64, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_evaluate.argtypes = [C.c_void_p, C.POINTER(OParams), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                      C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_systematic_resample.restype = C.c_uint64
        L.oracle_systematic_resample.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_void_p, C.c_uint64]
        L.oracle_residual_systematic_resample.restype = C.c_uint64
        L.oracle_residual_systematic_resample.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_void_p, C.c_uint64]
        L.oracle_residual_resample.restype = C.c_uint64
        L.oracle_residual_resample.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]
        L.oracle_wheel_resample.restype = None
        L.oracle_wheel_resample.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_metropolis_resample.restype = None
        L.oracle_metropolis_resample.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_rejection_resample.restype = None
        L.oracle_rejection_resample.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_draws_create.restype = C.c_void_p
        L.oracle_draws_create.argtypes = [C.c_uint64, C.c_uint64]
        L.oracle_draws_destroy.argtypes = [C.c_void_p]
        L.oracle_draws_used.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.oracle_motion_model.argtypes = [C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_motion_apply.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.oracle_reduce_scan.restype = C.c_int64
        L.oracle_reduce_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_float, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_reduce_scan_centres.restype = C.c_int64
        L.oracle_reduce_scan_centres.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_uint32, C.c_void_p, C.c_void_p]

    # -- motion update --
    def motion_model(self, variant, inputs, time_diff, a, ref_pose=None):
        """(mean[6], sigma[6], ref_pose') of the six normal distributions of ParticleCloud::motionUpdate variant `variant`."""
        inp = np.zeros(4, dtype=np.float64)
        inp[:len(inputs)] = inputs
        av = np.ascontiguousarray(a, dtype=np.float32)
        mean, sigma = np.zeros(6), np.zeros(6)
        rp = None if ref_pose is None else np.array(ref_pose, dtype=np.float32)
        rc = self.lib.oracle_motion_model(int(variant), _fp(inp), C.c_float(time_diff), _fp(av), _fp(mean), _fp(sigma),
                                          _fp(rp) if rp is not None else None)
        if rc:
            raise ValueError("bad motion variant")
        return mean, sigma, rp

    def motion_apply(self, particles, draws):
        ps = np.array(particles, dtype=np.float32, copy=True, order="C")
        dr = np.ascontiguousarray(draws, dtype=np.float64)
        assert dr.shape == (ps.shape[0], 6)
        self.lib.oracle_motion_apply(_fp(ps), ps.shape[0], _fp(dr))
        return ps

    # -- scan reduction --
    def reduce_scan(self, points, ring, cell, n_rings=128, ring_desync=False):
        """(points_out [m,3], src_index [m]) or raises ValueError when a ring is outside [0, n_rings)."""
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        rg = np.ascontiguousarray(ring, dtype=np.int32)
        n = pts.shape[0]
        out = np.empty((max(n, 1), 3), dtype=np.float32)
        src = np.empty(max(n, 1), dtype=np.uint32)
        m = int(self.lib.oracle_reduce_scan(_fp(pts), _fp(rg), n, C.c_float(cell), n_rings, 1 if ring_desync else 0, _fp(out), _fp(src)))
        if m < 0:
            raise ValueError("ring outside [0, n_rings)")
        return out[:m].copy(), src[:m].copy()

    def reduce_scan_centres(self, points, ring, cell=0.064, n_rings=128):
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        rg = None if ring is None else np.ascontiguousarray(ring, dtype=np.int32)
        n = pts.shape[0]
        out = np.empty((max(n, 1), 3), dtype=np.float32)
        src = np.empty(max(n, 1), dtype=np.uint32)
        m = int(self.lib.oracle_reduce_scan_centres(_fp(pts), _fp(rg) if rg is not None else None, n, C.c_double(cell), n_rings,
                                                    _fp(out), _fp(src)))
        if m < 0:
            raise ValueError("ring outside [0, n_rings)")
        return out[:m].copy(), src[:m].copy()

    # -- map --
    def map_create(self, mn, mx, res, init):
        a = np.asarray(mn, dtype=np.float32)
        b = np.asarray(mx, dtype=np.float32)
        return self.lib.oracle_map_create(_fp(a), _fp(b), C.c_float(res), C.c_float(init))

    def map_from_arrays(self, coef, grid_occ, data):
        c = Coef()
        C.memmove(C.byref(c), C.byref(coef), C.sizeof(Coef))
        g = np.ascontiguousarray(grid_occ, dtype=np.int32)
        d = np.ascontiguousarray(data, dtype=np.float32)
        return self.lib.oracle_map_from_arrays(C.byref(c), _fp(g), _fp(d))

    def map_set_data(self, m, cells):
        cells = np.ascontiguousarray(cells, dtype=np.float32)
        return self.lib.oracle_map_set_data(m, _fp(cells), cells.shape[0])

    def map_arrays(self, m):
        c = m.contents.coef
        occ = np.ctypeslib.as_array(m.contents.grid_occ, shape=(int(c.grid_occ_size),)).copy()
        data = (np.ctypeslib.as_array(m.contents.data, shape=(int(c.data_size),)).copy() if c.data_size
                else np.zeros(0, dtype=np.float32))
        return c, occ, data

    def map_destroy(self, m):
        self.lib.oracle_map_destroy(m)

    def get_entries(self, m, xyz, mode):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        out = np.empty(xyz.shape[0], dtype=np.float32)
        self.lib.oracle_get_entries(m, _fp(xyz), xyz.shape[0], mode, _fp(out))
        return out

    def get_index(self, m, x, y, z, mode):
        return int(self.lib.oracle_get_index(m, C.c_float(x), C.c_float(y), C.c_float(z), mode))

    def pose_matrices(self, particles, tf):
        particles = np.ascontiguousarray(particles, dtype=np.float32)
        tf = np.ascontiguousarray(tf, dtype=np.float32)
        out = np.empty((particles.shape[0], 12), dtype=np.float32)
        for i in range(particles.shape[0]):
            pose = np.ascontiguousarray(particles[i, :6])
            row = np.empty(12, dtype=np.float32)
            self.lib.oracle_pose_matrix(_fp(pose), _fp(tf), _fp(row))
            out[i] = row
        return out

    def evaluate(self, m, params, particles, points, tf, mode=NEG_AS_MISS, want_idx=False, want_hits=True):
        """Returns dict(status, normalised particles, raw, mean, idx, hits, weight_sum)."""
        ps = np.array(particles, dtype=np.float32, copy=True, order="C")
        pts = np.ascontiguousarray(points, dtype=np.float32)
        tf = np.ascontiguousarray(tf, dtype=np.float32)
        n, p = ps.shape[0], pts.shape[0]
        raw = np.empty(n, dtype=np.float32)
        mean = np.zeros(6, dtype=np.float32)
        idx = np.empty((n, p), dtype=np.uint32) if want_idx else None
        hits = np.empty(n, dtype=np.uint32) if want_hits else None
        wsum = C.c_float(0)
        prm = OParams(*params)
        rc = self.lib.oracle_evaluate(m, C.byref(prm), _fp(ps), n, _fp(pts), p, _fp(tf), mode, _fp(raw), _fp(mean),
                                      _fp(idx) if want_idx else None, _fp(hits) if want_hits else None, C.byref(wsum))
        return dict(status=rc, particles=ps, raw=raw, mean=mean, idx=idx, hits=hits, weight_sum=wsum.value)

    def pose_weight64(self, m, params, mat12, points, mode=NEG_AS_MISS):
        pts = np.ascontiguousarray(points, dtype=np.float32)
        mat = np.ascontiguousarray(mat12, dtype=np.float32)
        w64 = C.c_double(0)
        prm = OParams(*params)
        w32 = self.lib.oracle_pose_weight(m, C.byref(prm), _fp(mat), _fp(pts), pts.shape[0], mode, None, None, C.byref(w64))
        return float(w32), w64.value

    def systematic_resample(self, weights, u0, cap=None):
        w = np.ascontiguousarray(weights, dtype=np.float32)
        n = w.shape[0]
        cap = cap or (n + n // 8 + 64)
        parents = np.empty(cap, dtype=np.uint32)
        m = int(self.lib.oracle_systematic_resample(_fp(w), n, C.c_float(u0), _fp(parents), cap))
        return m, parents[:min(m, cap)]


    def residual_systematic_resample(self, weights, u0, cap=None):
        w = np.ascontiguousarray(weights, dtype=np.float32)
        n = w.shape[0]
        cap = cap or (2 * n + 64)
        parents = np.empty(cap, dtype=np.uint32)
        m = int(self.lib.oracle_residual_systematic_resample(_fp(w), n, C.c_float(u0), _fp(parents), cap))
        return m, parents[:min(m, cap)]

    def residual_resample(self, weights, draws):
        """(output length, parents, draws consumed) of the Residual resampler fed the given index draws."""
        w = np.ascontiguousarray(weights, dtype=np.float32)
        d = np.ascontiguousarray(draws, dtype=np.uint64)
        n = w.shape[0]
        parents = np.empty(n, dtype=np.uint32)
        used = C.c_uint64(0)
        m = int(self.lib.oracle_residual_resample(_fp(w), n, _fp(d), d.shape[0], _fp(parents), C.byref(used)))
        return m, parents[:m], int(used.value)


    def drawn_resample(self, method, weights, draws, steps=50):
        """Parents of the Wheel (3) / Metropolis (4) / Rejection (5) resampler restatements fed the draws of a NativeDraws."""
        w = np.ascontiguousarray(weights, dtype=np.float32)
        n = w.shape[0]
        parents = np.empty(n, dtype=np.uint32)
        if method == 3:
            self.lib.oracle_wheel_resample(_fp(w), n, draws.real_wheel_ptr, draws.handle, _fp(parents))
        elif method == 4:
            self.lib.oracle_metropolis_resample(_fp(w), n, steps, draws.real_ptr, draws.index_ptr, draws.handle, _fp(parents))
        elif method == 5:
            self.lib.oracle_rejection_resample(_fp(w), n, draws.real_ptr, draws.index_ptr, draws.handle, _fp(parents))
        else:
            raise ValueError(method)
        return parents

    def draws(self, seed, n):
        """The oracle's own deterministic draw source (splitmix64), usable where oracle/_ref is not built."""
        return NativeDraws(self.lib, "oracle", seed, n)


class NativeDraws:
    """A seeded native draw source behind C callbacks (`user` = handle): the oracle's splitmix64 ("oracle") or the reference's
    std::mt19937 with its own distribution objects ("ref", oracle/ref_harness.cpp DrawSource). Feed two equally seeded
    instances to two implementations to give them the same draws in the same interleaving."""

    def __init__(self, lib, prefix, seed, n):
        self._lib, self._prefix = lib, prefix
        create = getattr(lib, f"{prefix}_draws_create")
        create.restype = C.c_void_p
        create.argtypes = [C.c_uint32 if prefix == "ref" else C.c_uint64, C.c_uint64]
        getattr(lib, f"{prefix}_draws_destroy").argtypes = [C.c_void_p]
        getattr(lib, f"{prefix}_draws_used").argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        self.handle = C.c_void_p(create(seed, n))
        self.real_fn = getattr(lib, f"{prefix}_draw_real")
        self.real_wheel_fn = getattr(lib, f"{prefix}_draw_real_wheel") if prefix == "ref" else self.real_fn
        self.index_fn = getattr(lib, f"{prefix}_draw_index")
        self.real_ptr = C.cast(self.real_fn, C.c_void_p)
        self.real_wheel_ptr = C.cast(self.real_wheel_fn, C.c_void_p)
        self.index_ptr = C.cast(self.index_fn, C.c_void_p)

    def used(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        getattr(self._lib, f"{self._prefix}_draws_used")(self.handle, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def source(self):
        """As the product's DrawSource (tsdf_localization_b200.evaluator)."""
        from tsdf_localization_b200 import DrawSource
        return DrawSource(real=self.real_fn, index=self.index_fn, user=self.handle, wheel_real=self.real_wheel_fn)

    def close(self):
        if self.handle:
            getattr(self._lib, f"{self._prefix}_draws_destroy")(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ref_lib_path(threads: int | None = None) -> Path:
    name = "libtsdf_ref.so" if threads is None else f"libtsdf_ref_t{threads}.so"
    return ORACLE_DIR / "_ref" / name


def ref_available() -> bool:
    return ref_lib_path().exists()


def ref_shim_path() -> Path:
    return ORACLE_DIR / "_ref" / "libtsdf_ref_shim.so"


def ref_cuda_path() -> Path:
    """The reference's OWN CUDA evaluator (src/cuda/*.cu compiled unmodified for sm_100a) behind the same harness."""
    return ORACLE_DIR / "_ref" / "libtsdf_ref_cuda.so"


class Ref:
    """The verbatim reference (oracle/ref_harness.cpp)."""

    def __init__(self, threads: int | None = None, shim: bool = False, cuda: bool = False):
        """shim=True: the same reference classes linked against the product's drop-in CudaEvaluator shim + libtsdfloc.so
        (evaluate(use_cuda=True) and GpuSystematicResampler then run on the B200).
        cuda=True: linked against the reference's own CUDA evaluator (evaluate(use_cuda=True) runs the reference's kernels)."""
        self.lib = C.CDLL(str(ref_cuda_path() if cuda else ref_shim_path() if shim else ref_lib_path(threads)))
        L = self.lib
        if shim:
            L.ref_gpu_systematic_resample.restype = C.c_uint64
            L.ref_gpu_systematic_resample.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64]
            L.ref_gpu_resample_method.restype = C.c_uint64
            L.ref_gpu_resample_method.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64]
        L.ref_last_error.restype = C.c_char_p
        L.ref_omp_threads.restype = C.c_uint
        L.ref_map_create.restype = C.c_void_p
        L.ref_map_create.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float]
        L.ref_map_destroy.argtypes = [C.c_void_p]
        L.ref_map_set_data.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.ref_map_adopt_arrays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.ref_map_get_coef.argtypes = [C.c_void_p, C.POINTER(Coef)]
        L.ref_map_grid_occ.restype = C.POINTER(C.c_int32)
        L.ref_map_grid_occ.argtypes = [C.c_void_p]
        L.ref_map_data.restype = C.POINTER(C.c_float)
        L.ref_map_data.argtypes = [C.c_void_p]
        L.ref_map_get_entries.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_create_tsdf_map.restype = C.c_void_p
        L.ref_create_tsdf_map.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_float, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.ref_eval_create.restype = C.c_void_p
        L.ref_eval_create.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
        L.ref_eval_destroy.argtypes = [C.c_void_p]
        L.ref_eval_create_cell.restype = C.c_void_p
        L.ref_eval_create_cell.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
        L.ref_evaluate_cloud.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int,
                                         C.c_uint32, C.c_void_p, C.POINTER(C.c_uint64)]
        L.ref_evaluate.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_pose_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_systematic_resample.restype = C.c_uint64
        L.ref_systematic_resample.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_resample_method.restype = C.c_uint64
        L.ref_resample_method.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_uniform_index_draws.restype = None
        L.ref_uniform_index_draws.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]

        if not shim:
            L.ref_reduce_scan.restype = C.c_int64
            L.ref_reduce_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_float, C.c_void_p, C.c_uint64]

    def omp_threads(self):
        return int(self.lib.ref_omp_threads())

    def reduce_scan(self, points, ring, cell):
        """TSDFEvaluator::evaluateParticles' own reduction (verbatim); rings must be < 64."""
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        rg = np.ascontiguousarray(ring, dtype=np.int16)
        n = pts.shape[0]
        out = np.empty((max(n, 1), 3), dtype=np.float32)
        m = int(self.lib.ref_reduce_scan(_fp(pts), _fp(rg), n, C.c_float(cell), _fp(out), n))
        if m < 0:
            raise RuntimeError(self.last_error())
        return out[:m].copy()

    def map_create(self, mn, mx, res, init):
        a = np.asarray(mn, dtype=np.float32)
        b = np.asarray(mx, dtype=np.float32)
        return self.lib.ref_map_create(_fp(a), _fp(b), C.c_float(res), C.c_float(init))

    def map_set_data(self, m, cells):
        cells = np.ascontiguousarray(cells, dtype=np.float32)
        return self.lib.ref_map_set_data(m, _fp(cells), cells.shape[0])

    def map_adopt_arrays(self, m, grid_occ, data):
        """Hand the reference's map the arrays setData would have built (large synthetic maps)."""
        g = np.ascontiguousarray(grid_occ, dtype=np.int32)
        d = np.ascontiguousarray(data, dtype=np.float32)
        return self.lib.ref_map_adopt_arrays(m, _fp(g), _fp(d), d.shape[0])

    def map_coef(self, m):
        c = Coef()
        self.lib.ref_map_get_coef(m, C.byref(c))
        return c

    def map_arrays(self, m):
        c = Coef()
        self.lib.ref_map_get_coef(m, C.byref(c))
        occ = np.ctypeslib.as_array(self.lib.ref_map_grid_occ(m), shape=(int(c.grid_occ_size),)).copy()
        data = (np.ctypeslib.as_array(self.lib.ref_map_data(m), shape=(int(c.data_size),)).copy() if c.data_size
                else np.zeros(0, dtype=np.float32))
        return c, occ, data

    def map_destroy(self, m):
        self.lib.ref_map_destroy(m)

    def mcl_write(self, name, points, rings, particles, tf, pose7):
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        rg = np.ascontiguousarray(rings, dtype=np.int32)
        ps = np.ascontiguousarray(particles, dtype=np.float32).reshape(-1, 7)
        t = np.ascontiguousarray(tf, dtype=np.float32)
        po = np.ascontiguousarray(pose7, dtype=np.float32)
        self.lib.ref_mcl_write.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        if self.lib.ref_mcl_write(str(name).encode(), _fp(pts), _fp(rg), len(pts), _fp(ps), len(ps), _fp(t), _fp(po)):
            raise RuntimeError(self.last_error())

    def mcl_read(self, name):
        self.lib.ref_mcl_read.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
        p, n = C.c_uint64(0), C.c_uint64(0)
        if self.lib.ref_mcl_read(str(name).encode(), C.byref(p), C.byref(n), None, None, None, None, None):
            raise RuntimeError(self.last_error())
        pts = np.empty((p.value, 3), dtype=np.float32)
        rg = np.empty(p.value, dtype=np.int32)
        ps = np.empty((n.value, 7), dtype=np.float32)
        t = np.empty(16, dtype=np.float32)
        po = np.empty(7, dtype=np.float32)
        if self.lib.ref_mcl_read(str(name).encode(), C.byref(p), C.byref(n), _fp(pts), _fp(rg), _fp(ps), _fp(t), _fp(po)):
            raise RuntimeError(self.last_error())
        return pts, rg, ps, t, po

    def create_tsdf_map(self, chunk_pos, chunk_data, sigma=0.1):
        """createTSDFMap (map_util.h:17-154), verbatim, on in-memory chunks. Returns (map handle, free_map [n, 3])."""
        pos = np.ascontiguousarray(chunk_pos, dtype=np.int32).reshape(-1, 3)
        dat = np.ascontiguousarray(chunk_data, dtype=np.uint32).reshape(len(pos), -1)
        cap = dat.size
        free = np.empty((max(cap, 1), 3), dtype=np.float32)
        n_free = C.c_uint64(0)
        h = self.lib.ref_create_tsdf_map(_fp(pos), _fp(dat), len(pos), C.c_float(sigma), _fp(free), cap, C.byref(n_free))
        if not h:
            raise RuntimeError(self.last_error())
        return h, free[:int(n_free.value)].copy()

    def get_entries(self, m, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        out = np.empty(xyz.shape[0], dtype=np.float32)
        self.lib.ref_map_get_entries(m, _fp(xyz), xyz.shape[0], _fp(out))
        return out

    def eval_create(self, m, a_hit=0.9, a_range=0.1, a_max=0.0, max_range=100.0):
        return self.lib.ref_eval_create(m, a_hit, a_range, a_max, max_range)

    def eval_create_cell(self, m, cell, a_hit=0.9, a_range=0.1, a_max=0.0, max_range=100.0):
        return self.lib.ref_eval_create_cell(m, a_hit, a_range, a_max, max_range, cell)

    def evaluate_cloud(self, e, particles, points, ring, use_cuda=False, desync=False, n_rings=64):
        """TSDFEvaluator::evaluateParticles(..., use_cuda, ignore_tf=true) on a packed cloud. Returns (rc, particles with the
        normalised weights, pose[7], error text, reduced scan size (B200 shim build with use_cuda only))."""
        ps = np.array(particles, dtype=np.float32, copy=True, order="C")
        pts = np.ascontiguousarray(points, dtype=np.float32)
        rg = np.ascontiguousarray(ring, dtype=np.int16)
        pose = np.zeros(7, dtype=np.float64)
        used = C.c_uint64(0)
        rc = self.lib.ref_evaluate_cloud(e, _fp(ps), ps.shape[0], _fp(pts), _fp(rg), pts.shape[0], 1 if use_cuda else 0,
                                         1 if desync else 0, n_rings, _fp(pose), C.byref(used))
        err = self.lib.ref_last_error().decode() if rc else ""
        return rc, ps, pose, err, int(used.value)

    def eval_destroy(self, e):
        self.lib.ref_eval_destroy(e)

    def last_error(self) -> str:
        return self.lib.ref_last_error().decode()

    def gpu_systematic_resample(self, particles, seed, cap=None):
        ps = np.ascontiguousarray(particles, dtype=np.float32)
        n = ps.shape[0]
        cap = cap or (n + n // 8 + 64)
        out = np.empty((cap, 7), dtype=np.float32)
        m = int(self.lib.ref_gpu_systematic_resample(_fp(ps), n, seed, _fp(out), cap))
        if m == 2 ** 64 - 1:
            raise RuntimeError(self.last_error())
        return m, out[:min(m, cap)]

    def evaluate(self, e, particles, points, tf, use_cuda: bool = False):
        ps = np.array(particles, dtype=np.float32, copy=True, order="C")
        pts = np.ascontiguousarray(points, dtype=np.float32)
        tf = np.ascontiguousarray(tf, dtype=np.float32)
        pose = np.zeros(7, dtype=np.float64)
        rc = self.lib.ref_evaluate(e, _fp(ps), ps.shape[0], _fp(pts), pts.shape[0], _fp(tf), 1 if use_cuda else 0, _fp(pose))
        err = self.lib.ref_last_error().decode() if rc else ""
        return rc, ps, pose, err

    def pose_weights(self, e, mats12, points):
        mats = np.ascontiguousarray(mats12, dtype=np.float32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.empty(mats.shape[0], dtype=np.float32)
        self.lib.ref_pose_weights(e, _fp(mats), mats.shape[0], _fp(pts), pts.shape[0], _fp(out))
        return out

    def resample_method(self, method, particles, seed, cap=None):
        """The verbatim ResidualResampler (method 1) / ResidualSystematicResampler (2) / WheelResampler (3) / MetropolisResampler
        (4, steps via set_metropolis_steps) / RejectionResampler (5) with a seeded generator: (length, particles out, the
        uniform(0,1) draw of method 2)."""
        ps = np.ascontiguousarray(particles, dtype=np.float32)
        n = ps.shape[0]
        cap = cap or (2 * n + 64)
        out = np.empty((cap, 7), dtype=np.float32)
        u = C.c_float(0)
        m = int(self.lib.ref_resample_method(method, _fp(ps), n, seed, _fp(out), cap, C.byref(u)))
        return m, out[:min(m, cap)], u.value

    def gpu_resample_method(self, method, particles, seed, cap=None):
        ps = np.ascontiguousarray(particles, dtype=np.float32)
        n = ps.shape[0]
        cap = cap or (2 * n + 64)
        out = np.empty((cap, 7), dtype=np.float32)
        m = int(self.lib.ref_gpu_resample_method(method, _fp(ps), n, seed, _fp(out), cap))
        if m == 2 ** 64 - 1:
            raise RuntimeError(self.last_error())
        return m, out[:min(m, cap)]

    def set_metropolis_steps(self, steps):
        self.lib.ref_set_metropolis_steps.argtypes = [C.c_uint64]
        self.lib.ref_set_metropolis_steps(steps)

    def draws(self, seed, n):
        """std::mt19937(seed) with the reference's distribution objects, as callbacks."""
        return NativeDraws(self.lib, "ref", seed, n)

    def uniform_index_draws(self, seed, n, count):
        out = np.empty(count, dtype=np.uint64)
        self.lib.ref_uniform_index_draws(seed, n, count, _fp(out))
        return out

    def systematic_resample(self, particles, seed, cap=None):
        ps = np.ascontiguousarray(particles, dtype=np.float32)
        n = ps.shape[0]
        cap = cap or (n + n // 8 + 64)
        out = np.empty((cap, 7), dtype=np.float32)
        u0 = C.c_float(0)
        m = int(self.lib.ref_systematic_resample(_fp(ps), n, seed, _fp(out), cap, C.byref(u0)))
        return m, out[:min(m, cap)], u0.value


def ref_pc_path() -> Path:
    return ORACLE_DIR / "_ref" / "libtsdf_ref_pc.so"


class RefPC:
    """The verbatim reference ParticleCloud motion update (oracle/pc_harness.cpp)."""

    def __init__(self):
        self.lib = C.CDLL(str(ref_pc_path()))
        self.lib.ref_pc_motion_update.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
        self.lib.ref_pc_draws.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]

    def motion_update(self, variant, inputs, dt, a, seed, particles, ref_pose=None):
        inp = np.zeros(4, dtype=np.float64)
        inp[:len(inputs)] = inputs
        av = np.ascontiguousarray(a, dtype=np.float32)
        ps = np.array(particles, dtype=np.float32, copy=True, order="C")
        rp = np.zeros(6, dtype=np.float32) if ref_pose is None else np.array(ref_pose, dtype=np.float32)
        rc = self.lib.ref_pc_motion_update(int(variant), _fp(inp), C.c_double(dt), _fp(av), seed, _fp(ps), ps.shape[0], _fp(rp))
        assert rc == 0
        return ps, rp

    def draws(self, seed, mean, sigma, n):
        m = np.ascontiguousarray(mean, dtype=np.float64)
        s = np.ascontiguousarray(sigma, dtype=np.float64)
        out = np.empty((n, 6), dtype=np.float64)
        self.lib.ref_pc_draws(seed, _fp(m), _fp(s), n, _fp(out))
        return out
