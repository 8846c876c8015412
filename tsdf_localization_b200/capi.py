"""ctypes binding of include/tsdfloc.h. Fails loudly when the CUDA library is missing — there is no fallback."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parent

OK, E_BAD_ARG, E_CUDA, E_NO_VALID_PARTICLE, E_EMPTY_SCAN, E_CAPACITY, E_STATE = range(7)


REDUCE_RING_DESYNC_LIKE_REFERENCE, REDUCE_EMIT_CENTRES = 1, 2
MOTION_NOISE, MOTION_ODOM, MOTION_IMU, MOTION_NOISE_IMU = range(4)
INIT_NORMAL, INIT_UNIFORM, INIT_FREE_MAP = range(3)
NEG_MISS, NEG_SATURATE_LIKE_REF_GPU = range(2)
TUNE_SPATIAL_ORDER, TUNE_EVAL_PAIRING, TUNE_DIVISION, TUNE_STAGE_TIMERS, TUNE_EVAL_REGISTERS, TUNE_GRAPHS, TUNE_EVAL_CHUNKS = range(7)
DIV_IEEE, DIV_THREE, DIV_BRACKET = range(3)
RESAMPLE_SYSTEMATIC, RESAMPLE_RESIDUAL, RESAMPLE_RESIDUAL_SYSTEMATIC = range(3)
RESAMPLE_WHEEL, RESAMPLE_METROPOLIS, RESAMPLE_REJECTION = 3, 4, 5
INDEX_DRAW_FN = C.CFUNCTYPE(C.c_uint64, C.c_void_p)
REAL_DRAW_FN = C.CFUNCTYPE(C.c_float, C.c_void_p)


class Draws(C.Structure):
    """tsdfloc_draws: the random draws of the Wheel / Metropolis / Rejection resamplers, as callbacks."""
    _fields_ = [("real", REAL_DRAW_FN), ("index", INDEX_DRAW_FN), ("user", C.c_void_p), ("metropolis_steps", C.c_uint64),
                ("max_draws", C.c_uint64)]


class LibraryNotBuilt(RuntimeError):
    pass


class TsdflocError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


class MapDesc(C.Structure):
    """tsdfloc_map_desc == CudaSubVoxelMap::MapCoef (cuda_sub_voxel_map.h:22-50)."""
    _fields_ = [
        ("dim", C.c_uint64 * 3), ("min", C.c_float * 3), ("max", C.c_float * 3), ("resolution", C.c_float),
        ("init_value", C.c_float), ("up_dim", C.c_uint64 * 3), ("up_dim_2", C.c_uint64), ("sub_dim", C.c_uint64),
        ("sub_dim_2", C.c_uint64), ("grid_occ_size", C.c_uint64), ("data_size", C.c_uint64),
    ]


class Params(C.Structure):
    _fields_ = [("a_hit", C.c_float), ("a_range", C.c_float), ("a_max", C.c_float), ("max_range", C.c_float),
                ("per_point", C.c_int32), ("neg_policy", C.c_int32)]


def lib_path() -> Path:
    # TSDFLOC_LIB: alternative build of the same library (kernel tuning experiments); never a different implementation
    override = os.environ.get("TSDFLOC_LIB")
    return Path(override) if override else PKG / "lib" / "libtsdfloc.so"


_LIB = None

_vp, _u64, _u32p, _fp, _i32p = C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.POINTER(C.c_int32)

# name -> (restype, argtypes): every symbol include/tsdfloc.h declares
SIGNATURES = {
    "tsdfloc_default_params": (None, [C.POINTER(Params)]),
    "tsdfloc_abi_version": (C.c_int, []),
    "tsdfloc_status_string": (C.c_char_p, [C.c_int]),
    "tsdfloc_last_error": (C.c_char_p, [_vp]),
    "tsdfloc_map_create": (C.c_int, [_fp, _fp, C.c_float, C.c_float, C.POINTER(_vp)]),
    "tsdfloc_map_set_data": (C.c_int, [_vp, _fp, _u64]),
    "tsdfloc_map_get_desc": (C.POINTER(MapDesc), [_vp]),
    "tsdfloc_map_grid_occ": (_i32p, [_vp]),
    "tsdfloc_map_data": (_fp, [_vp]),
    "tsdfloc_map_destroy": (None, [_vp]),
    "tsdfloc_map_from_chunks": (C.c_int, [_vp, _vp, _u64, C.c_float, C.POINTER(_vp)]),
    "tsdfloc_map_from_chunks_gpu": (C.c_int, [_vp, _vp, _u64, C.c_float, C.c_int, C.POINTER(_vp)]),
    "tsdfloc_map_free_points": (_fp, [_vp, C.POINTER(_u64)]),
    "tsdfloc_create_from_chunks": (C.c_int, [_vp, _vp, _u64, C.c_float, C.POINTER(Params), C.c_int, C.POINTER(_vp)]),
    "tsdfloc_map_desc_of": (C.c_int, [_vp, C.POINTER(MapDesc)]),
    "tsdfloc_free_map_device": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_u64)]),
    "tsdfloc_mcl_read": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "tsdfloc_mcl_free": (None, [_vp]),
    "tsdfloc_mcl_n_points": (_u64, [_vp]),
    "tsdfloc_mcl_n_particles": (_u64, [_vp]),
    "tsdfloc_mcl_points": (_fp, [_vp]),
    "tsdfloc_mcl_rings": (_i32p, [_vp]),
    "tsdfloc_mcl_particles": (_fp, [_vp]),
    "tsdfloc_mcl_tf": (_fp, [_vp]),
    "tsdfloc_mcl_pose": (_fp, [_vp]),
    "tsdfloc_mcl_write": (C.c_int, [C.c_char_p, _vp, _vp, _u64, _vp, _u64, _fp, _fp]),
    "tsdfloc_likelihood_value": (C.c_float, [C.c_float, C.c_float]),
    "tsdfloc_likelihood_init": (C.c_float, [C.c_float]),
    "tsdfloc_create": (C.c_int, [C.POINTER(MapDesc), _vp, _vp, C.POINTER(Params), C.c_int, C.POINTER(_vp)]),
    "tsdfloc_destroy": (None, [_vp]),
    "tsdfloc_sensor_update": (C.c_int, [_vp, _vp, _u64, _vp, _u64, _fp, _fp]),
    "tsdfloc_resample_systematic": (C.c_int, [_vp, C.c_float, _vp, _u64, C.POINTER(_u64), _vp]),
    "tsdfloc_resample_particles": (C.c_int, [_vp, _vp, _u64, C.c_float, _vp, _u64, C.POINTER(_u64), _vp]),
    "tsdfloc_residual_systematic_counts": (C.c_int, [_vp, _u64, _u64, C.c_float, _vp, C.POINTER(_u64)]),
    "tsdfloc_residual_runs": (C.c_int, [_vp, _u64, _u64, INDEX_DRAW_FN, _vp, _u64, _vp, _vp, _u64, C.POINTER(_u64), C.POINTER(_u64)]),
    "tsdfloc_resample_expand": (C.c_int, [_vp, _vp, _vp, _u64, _vp, _u64, C.POINTER(_u64), _vp]),
    "tsdfloc_resample_expand_device": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _u64, _u64, _vp, C.POINTER(_vp), C.c_uint32, _vp, _vp]),
    "tsdfloc_resample": (C.c_int, [_vp, C.c_int, _vp, _u64, C.c_float, INDEX_DRAW_FN, _vp, _vp, _u64, C.POINTER(_u64), _vp]),
    "tsdfloc_wheel_parents": (C.c_int, [_vp, _u64, _u64, REAL_DRAW_FN, _vp, _vp]),
    "tsdfloc_metropolis_parents": (C.c_int, [_vp, _u64, _u64, _u64, REAL_DRAW_FN, INDEX_DRAW_FN, _vp, _vp]),
    "tsdfloc_rejection_parents": (C.c_int, [_vp, _u64, _u64, REAL_DRAW_FN, INDEX_DRAW_FN, _vp, _u64, _vp, C.POINTER(_u64)]),
    "tsdfloc_resample_drawn": (C.c_int, [_vp, C.c_int, _vp, _u64, C.POINTER(Draws), _vp, _u64, C.POINTER(_u64), _vp]),
    "tsdfloc_cdf_device": (C.c_int, [_vp, _vp, _u64, _vp, _vp]),
    "tsdfloc_debug_eval": (C.c_int, [_vp, _vp, _u64, _vp, _u64, _fp, _vp, _vp, _vp]),
    "tsdfloc_set_scan_device": (C.c_int, [_vp, _vp, _u64, _vp]),
    "tsdfloc_set_scan_host": (C.c_int, [_vp, _vp, _u64, _vp]),
    "tsdfloc_eval_device": (C.c_int, [_vp, _vp, _u64, _u64, _u64, _fp, _vp, _vp]),
    "tsdfloc_eval_device_peers": (C.c_int, [_vp, _vp, _u64, _u64, _u64, _fp, _vp, C.POINTER(_vp), C.c_uint32, _vp]),
    "tsdfloc_draw_device_peers": (C.c_int, [_vp, _vp, _u64, C.c_float, _u64, _u64, _vp, C.POINTER(_vp), C.c_uint32, _vp, _vp]),
    "tsdfloc_multi_create": (C.c_int, [C.POINTER(MapDesc), _vp, _vp, C.POINTER(Params), C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    "tsdfloc_multi_destroy": (None, [_vp]),
    "tsdfloc_multi_device_count": (C.c_int, [_vp]),
    "tsdfloc_multi_last_error": (C.c_char_p, [_vp]),
    "tsdfloc_multi_ctx": (_vp, [_vp, C.c_int]),
    "tsdfloc_multi_sensor_update": (C.c_int, [_vp, _vp, _u64, _vp, _u64, _fp, _fp]),
    "tsdfloc_multi_resample_systematic": (C.c_int, [_vp, C.c_float, _vp, _u64, C.POINTER(_u64)]),
    "tsdfloc_multi_sensor_update_cloud": (C.c_int, [_vp, _vp, _u64, _vp, _u64, _vp, _u64, C.c_int, _u64, C.c_double, C.c_uint32, C.c_uint32,
                                                    _fp, _fp, C.POINTER(_u64)]),
    "tsdfloc_normalize_device": (C.c_int, [_vp, _vp, _u64, _vp, _vp, _vp]),
    "tsdfloc_draw_device": (C.c_int, [_vp, _vp, _u64, C.c_float, _u64, _u64, _vp, _vp, _vp]),
    "tsdfloc_check": (C.c_int, [_vp, C.POINTER(_u64), C.POINTER(C.c_double), _vp]),
    "tsdfloc_update_device": (C.c_int, [_vp, _vp, _u64, _vp, _u64, _fp, C.c_float, _vp, _u64, _vp, _vp]),
    "tsdfloc_reduce_scan_device": (C.c_int, [_vp, _vp, _vp, _u64, C.c_double, C.c_uint32, C.c_uint32, _vp, _vp, _vp]),
    "tsdfloc_reduce_result": (C.c_int, [_vp, C.POINTER(_u64), _vp]),
    "tsdfloc_reduce_scan": (C.c_int, [_vp, _vp, _u64, _vp, _u64, C.c_int, _u64, C.c_double, C.c_uint32, C.c_uint32, _vp, _vp, _u64,
                                      C.POINTER(_u64)]),
    "tsdfloc_sensor_update_cloud": (C.c_int, [_vp, _vp, _u64, _vp, _u64, _vp, _u64, C.c_int, _u64, C.c_double, C.c_uint32, C.c_uint32,
                                              _fp, _fp, C.POINTER(_u64)]),
    "tsdfloc_motion_model": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.c_float, _fp, C.POINTER(C.c_double), C.POINTER(C.c_double), _fp]),
    "tsdfloc_motion_update_device": (C.c_int, [_vp, _vp, _u64, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp, _u64, _u64, _vp]),
    "tsdfloc_motion_update": (C.c_int, [_vp, _vp, _u64, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp, _u64, _u64]),
    "tsdfloc_init_particles_device": (C.c_int, [_vp, _vp, _u64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp, _u64, _u64, _u64, _vp]),
    "tsdfloc_init_particles": (C.c_int, [_vp, _vp, _u64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp, _u64, _u64, _u64]),
    "tsdfloc_best_particle": (C.c_int, [_vp, C.POINTER(C.c_int64), _fp, _fp, _vp]),
    "tsdfloc_host_u_sequence": (_u64, [C.c_float, _u64, C.c_double, _vp, _u64, _u32p, _u32p]),
    "tsdfloc_probe_gather": (C.c_int, [_vp, _u64, C.c_uint32, C.c_uint32, _fp, C.POINTER(_u64)]),
    "tsdfloc_eval_stats": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "tsdfloc_tune": (C.c_int, [_vp, C.c_int, C.c_int]),
    "tsdfloc_stage_times": (C.c_int, [_vp, _fp]),
    "tsdfloc_graph_stats": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_char_p)]),
    "tsdfloc_last_cdf_was_exact": (C.c_int, [_vp]),
    "tsdfloc_division_mode": (C.c_int, [_vp, C.POINTER(_u64)]),
    "tsdfloc_last_eval_ms": (C.c_int, [_vp, _fp]),
    "tsdfloc_kernel_launches": (_u64, [_vp]),
}


def load_library():
    """dlopen lib/libtsdfloc.so and type every entry point. Raises LibraryNotBuilt if it is not there."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not path.exists():
        raise LibraryNotBuilt(
            f"{path} is missing: build it with `python -m tsdf_localization_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback for the sensor update.")
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(lib, ctx, status: int) -> None:
    if status == OK:
        return
    msg = lib.tsdfloc_last_error(ctx)
    text = msg.decode() if msg else ""
    if not text:
        text = lib.tsdfloc_status_string(status).decode()
    raise TsdflocError(status, text)
