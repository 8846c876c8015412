"""Chained scan chunks of the evaluation kernel (csrc/tsdfloc_eval.cuh, kChain): a particle's sequential fp32 sum is carried
from warp to warp through global memory, chunk by chunk — the weights must be the very bits of the whole-scan walk (and so of
the oracle / the reference, cuda_eval_particles.h:200-211), for every pairing, register budget and chunk count, launch after
launch (the ticket and the links re-arm themselves)."""
import numpy as np
import pytest

import common
from tsdf_localization_b200 import CudaEvaluator, capi, synthetic as syn

pytestmark = pytest.mark.gpu

GT = (0.4, -0.3, 1.2, 0.01, -0.02, 0.4)
ROOM = dict(room_lo=(-3.0, -2.5, 0.0), room_hi=(3.0, 2.5, 3.0))


def _raw_weights(ev, ps, pts):
    """Un-normalised weights of the production launch (no dump instantiation): through the device-pointer stage call."""
    import ctypes as C
    import torch
    lib = capi.load_library()
    d_ps = torch.from_numpy(ps).cuda()
    d_pts = torch.from_numpy(pts).cuda()
    d_raw = torch.zeros(len(ps), dtype=torch.float32, device="cuda")
    tf = (C.c_float * 16)(*[float(v) for v in syn.IDENTITY_TF])
    torch.cuda.synchronize()      # the uploads ran on torch's stream, the stage calls use the ctx's own
    capi.check(lib, ev.ctx, lib.tsdfloc_set_scan_device(ev.ctx, C.c_void_p(d_pts.data_ptr()), len(pts), None))
    capi.check(lib, ev.ctx, lib.tsdfloc_eval_device(ev.ctx, C.c_void_p(d_ps.data_ptr()), len(ps), 0, len(ps), tf, C.c_void_p(d_raw.data_ptr()), None))
    torch.cuda.synchronize()
    return d_raw.cpu().numpy()


def test_chunked_sums_equal_whole_scan_sums_and_the_oracle(oracle):
    _, m = common.box_room(small=True)
    om = common.oracle_map_of(oracle, m)
    ev = CudaEvaluator(m)
    pts, _ = syn.make_scan("os1-128", GT, n_points=20000, **ROOM)
    ps = syn.tracking_particles(601, GT, sigma_xy=0.15)      # odd: the last pair-warp holds one particle
    want = oracle.evaluate(om, common.DEFAULT_PARAMS, ps, pts, syn.IDENTITY_TF)["raw"]
    try:
        for registers in (1, 2):
            for pairing in (1, 2):
                ev.tune(capi.TUNE_EVAL_REGISTERS, registers)
                ev.tune(capi.TUNE_EVAL_PAIRING, pairing)
                stats = {}
                for chunks in (1, 2, 3, 5, 64):          # 64 is cut to >= 8 summation blocks per chunk by the library
                    ev.tune(capi.TUNE_EVAL_CHUNKS, chunks)
                    before = ev.eval_stats()
                    for rep in range(3):                  # the chain re-arms itself
                        got = _raw_weights(ev, ps, pts)
                        assert got.tobytes() == want.tobytes(), (registers, pairing, chunks, rep)
                    after = ev.eval_stats()
                    stats[chunks] = {k: after[k] - before[k] for k in after}
                # same blocks, same sequential folds, same ties, whoever carried the sum
                assert all(v == stats[1] for v in stats.values()), stats
    finally:
        ev.close()


def test_many_more_units_than_warp_slots():
    """24,000 one-warp units of four chunks each on ~2,400 slots: every (chunk, particle) unit finds its predecessor's sum."""
    _, m = common.box_room(small=True)
    ev = CudaEvaluator(m)
    pts, _ = syn.make_scan("os1-128", GT, n_points=8192, **ROOM)
    ps = syn.tracking_particles(12000, GT, sigma_xy=0.2)
    try:
        ev.tune(capi.TUNE_EVAL_CHUNKS, 1)
        want = _raw_weights(ev, ps, pts)
        for chunks in (4, 2):
            ev.tune(capi.TUNE_EVAL_CHUNKS, chunks)
            for rep in range(2):
                assert _raw_weights(ev, ps, pts).tobytes() == want.tobytes(), (chunks, rep)
    finally:
        ev.close()


def test_automatic_rule_on_an_eighth_of_c3_and_through_the_full_update():
    """8,192 particles x 131,072 points (the per-GPU slice of C3 on 8 GPUs) is where the rule switches chunks on: same weights,
    same resampled set as whole-scan walks — through the host-buffer update and the resampler."""
    _, m = common.box_room()
    ps0, pts, _ = common.config_c3(8192)
    res = {}
    for chunks in (1, 0):
        ev = CudaEvaluator(m)
        ev.tune(capi.TUNE_EVAL_CHUNKS, chunks)
        ps = ps0.copy()
        ev.evaluate(ps, pts, syn.IDENTITY_TF)
        out = ev.resample_systematic(0.37 / len(ps), capacity=len(ps) + 64)
        res[chunks] = (ps, out.copy(), ev.eval_stats())
        ev.close()
    assert res[0][0].tobytes() == res[1][0].tobytes()
    assert res[0][1].tobytes() == res[1][1].tobytes()
    assert res[0][2] == res[1][2]
