"""CPU-side checks of libtsdfloc.so: it loads, exports every symbol include/tsdfloc.h declares, its host-only pieces
(map builder, likelihood LUT, U-recurrence table) agree with the oracle / the plain loop, and every compute entry point
FAILS LOUDLY when no B200 is present (there is no CPU fallback). No GPU compute calls here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import common
from tsdf_localization_b200 import CudaEvaluator, CudaSubVoxelMap, capi, likelihood_init, likelihood_value
from tsdf_localization_b200 import synthetic as syn

ROOT = Path(__file__).resolve().parent.parent


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_exported_and_typed(lib):
    header = (ROOT / "include" / "tsdfloc.h").read_text()
    declared = set(re.findall(r"\b(tsdfloc_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/tsdfloc.h but not exported by libtsdfloc.so"
    assert declared == set(capi.SIGNATURES), "ctypes binding and header disagree"
    assert lib.tsdfloc_abi_version() == 3


def test_status_strings_and_defaults(lib):
    assert lib.tsdfloc_status_string(capi.E_NO_VALID_PARTICLE) == b"No particle is valid!"   # cuda_evaluator.cu:366-369
    assert lib.tsdfloc_status_string(capi.OK) == b"ok"
    p = capi.Params()
    lib.tsdfloc_default_params(C.byref(p))
    assert (p.a_hit, p.a_range, p.a_max, p.max_range) == pytest.approx((0.9, 0.1, 0.0, 100.0))   # util.h:13-18


def test_no_cpu_fallback_without_gpu():
    if _has_gpu():
        pytest.skip("a GPU is present")
    _, m = common.box_room(small=True)
    with pytest.raises(RuntimeError) as ei:
        CudaEvaluator(m)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_multi_gpu_evaluator_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from tsdf_localization_b200 import MultiGpuEvaluator
    _, m = common.box_room(small=True)
    with pytest.raises(RuntimeError) as ei:
        MultiGpuEvaluator(m, [0, 1])
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_header_is_plain_c_and_links(tmp_path):
    """include/tsdfloc.h must be usable from C (the drop-in boundary is a C ABI): compile a C99 translation unit against it with
    -pedantic, link it against libtsdfloc.so and run it (no GPU call)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        pytest.skip("no C compiler")
    src = tmp_path / "abi.c"
    src.write_text('#include "tsdfloc.h"\n'
                   "int main(void) { tsdfloc_params p; tsdfloc_map_desc d; (void)d; tsdfloc_default_params(&p);\n"
                   "  return (tsdfloc_abi_version() == TSDFLOC_ABI_VERSION && p.max_range == 100.0f) ? 0 : 1; }\n")
    exe = tmp_path / "abi"
    lib_dir = capi.lib_path().parent
    subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(lib_dir.parent.parent / "include"),
                    str(src), "-L", str(lib_dir), "-ltsdfloc", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_structs_match_header_layout():
    assert C.sizeof(capi.MapDesc) == 120     # MapCoef is 120 bytes
    assert C.sizeof(capi.Params) == 24


@pytest.mark.parametrize("res", [0.05, 0.064, 0.1, 0.2])
def test_host_map_builder_matches_oracle(oracle, res):
    mn, mx = (-3.3, -2.0, -0.5), (4.1, 2.7, 2.2)
    rng = np.random.default_rng(4)
    cells = np.empty((5000, 4), dtype=np.float32)
    cells[:, :3] = rng.uniform(np.asarray(mn) + 0.01, np.asarray(mx) - 0.01, size=(5000, 3))
    cells[:, 3] = rng.uniform(0.1, 60.0, size=5000)
    m = CudaSubVoxelMap(*mn, *mx, res, 0.0)
    m.setData(cells)
    om = oracle.map_create(mn, mx, res, 0.0)
    assert oracle.map_set_data(om, cells) == 0
    oc, oocc, odata = oracle.map_arrays(om)
    d = m.coef()
    for f, _ in capi.MapDesc._fields_:
        a, b = getattr(d, f), getattr(oc, f)
        assert (tuple(a) == tuple(b)) if hasattr(a, "__len__") else (a == b), f
    assert np.array_equal(m.rawGridOcc(), oocc)
    assert m.rawData().tobytes() == odata.tobytes()
    oracle.map_destroy(om)


def test_box_room_map_matches_oracle(oracle):
    spec, m = common.box_room(small=True)
    om = oracle.map_create(spec.min, spec.max, spec.resolution, spec.init_value)
    assert oracle.map_set_data(om, spec.cells) == 0
    _, oocc, odata = oracle.map_arrays(om)
    assert np.array_equal(m.rawGridOcc(), oocc) and m.rawData().tobytes() == odata.tobytes()
    oracle.map_destroy(om)


def test_map_rejects_cells_outside_bounds():
    m = CudaSubVoxelMap(0, 0, 0, 2, 2, 2, 0.1, 0.0)
    with pytest.raises(ValueError):
        m.setData(np.array([[-0.5, 1, 1, 1.0]], dtype=np.float32))
    with pytest.raises(ValueError):
        m.setData(np.array([[1, 1, 2.5, 1.0]], dtype=np.float32))


def test_likelihood_lut_matches_oracle(oracle):
    for sigma in (0.1, 0.05, 0.3):
        assert np.float32(likelihood_init(sigma)).tobytes() == np.float32(oracle.lib.oracle_likelihood_init(sigma)).tobytes()
        for mm in range(-599, 600, 7):
            a = np.float32(likelihood_value(float(mm), sigma))
            b = np.float32(oracle.lib.oracle_likelihood_value(float(mm), sigma))
            assert a.tobytes() == b.tobytes(), (mm, sigma)
    assert likelihood_init(0.1) == 0.0
    assert abs(likelihood_value(0.0, 0.1) - 63.49) < 0.01   # range [0, 63.49]


def _u_loop(u0, n, limit, cap):
    inv = 1.0 / n
    u = np.float32(u0)
    out = []
    while float(u) < limit and len(out) < cap:
        out.append(u)
        nxt = np.float32(float(u) + inv)       # float += double, rounded to fp32 each step (novel_resampling.h:61-64)
        if not nxt > u:
            break
        u = nxt
    return np.array(out, dtype=np.float32)


@pytest.mark.parametrize("n", [1, 2, 3, 500, 5000, 65536, 100000, 1000000, 1048576, 3000017])
@pytest.mark.parametrize("u0_frac", [0.0, 0.37, 0.999999])
def test_u_recurrence_table_matches_loop(lib, n, u0_frac):
    u0 = float(np.float32(u0_frac / n))
    if u0 >= 1.0 / n:
        u0 = float(np.nextafter(np.float32(u0), np.float32(0)))
    for limit in (1.0, 0.99993, 1.00004, 0.25):
        cap = 2 * n + 64
        out = np.empty(cap, dtype=np.float32)
        nseg, flags = C.c_uint32(0), C.c_uint32(0)
        cnt = lib.tsdfloc_host_u_sequence(C.c_float(u0), n, C.c_double(limit), out.ctypes.data_as(C.c_void_p), cap,
                                          C.byref(nseg), C.byref(flags))
        ref = _u_loop(u0, n, limit, cap)
        assert flags.value & 1 == 0, "segment table overflow"
        assert cnt == len(ref), (n, u0, limit, cnt, len(ref), nseg.value)
        assert out[:cnt].tobytes() == ref.tobytes()
        assert nseg.value <= 192
