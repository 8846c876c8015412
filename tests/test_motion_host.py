"""tsdfloc_motion_model (product host code, csrc/host_motion.cpp) against the oracle, which tests/test_motion_oracle.py pins
against the reference's ParticleCloud::motionUpdate. Bit-exact means, sigmas and reference pose. CPU only (no compute call)."""
import ctypes as C

import numpy as np
import pytest

from test_motion_oracle import A_DEFAULT, A_MIXED, CASES
from tsdf_localization_b200 import capi


@pytest.mark.parametrize("variant,inputs", CASES)
@pytest.mark.parametrize("a", [A_DEFAULT, A_MIXED], ids=["a_default", "a_mixed"])
@pytest.mark.parametrize("dt", [0.1, 0.05, 1.7])
def test_motion_model_matches_oracle(lib, oracle, variant, inputs, a, dt):
    rp0 = np.array([0.3, -0.2, 0.0, 0.0, 0.0, 0.4], dtype=np.float32)
    want_mean, want_sigma, want_rp = oracle.motion_model(variant, inputs, np.float32(dt), a, rp0)
    inp = (C.c_double * 4)(*(list(inputs) + [0.0] * (4 - len(inputs))))
    av = np.asarray(a, dtype=np.float32)
    mean, sigma = (C.c_double * 6)(), (C.c_double * 6)()
    rp = rp0.copy()
    rc = lib.tsdfloc_motion_model(variant, inp, C.c_float(dt), av.ctypes.data_as(C.POINTER(C.c_float)), mean, sigma,
                                  rp.ctypes.data_as(C.POINTER(C.c_float)))
    assert rc == capi.OK
    assert np.array(list(mean)).tobytes() == want_mean.tobytes()
    assert np.array(list(sigma)).tobytes() == want_sigma.tobytes()
    assert rp.tobytes() == want_rp.tobytes()


def test_motion_model_rejects_unknown_variant(lib):
    inp = (C.c_double * 4)()
    av = np.zeros(12, dtype=np.float32)
    mean, sigma = (C.c_double * 6)(), (C.c_double * 6)()
    assert lib.tsdfloc_motion_model(9, inp, C.c_float(0.1), av.ctypes.data_as(C.POINTER(C.c_float)), mean, sigma, None) == capi.E_BAD_ARG
