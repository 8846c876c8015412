"""Steady-state CUDA graphs (csrc/tsdfloc_graph.inc): a fixed-shape device-resident update (tsdfloc_update_device) is recorded
on its second call and replayed from then on — and every replay must produce exactly the bits of the kernel-by-kernel launches (TSDFLOC_TUNE_GRAPHS = 0), with a
different sensor transform and a different u0 on every call (the two things that are patched into the recorded kernel nodes).
The orchestration this replaces: src/cuda/cuda_evaluator.cu:118-428 (re-issued per scan)."""
import ctypes as C

import numpy as np
import pytest

import common
from tsdf_localization_b200 import CudaEvaluator, capi, synthetic as syn

pytestmark = pytest.mark.gpu

GT = (0.4, -0.3, 1.2, 0.01, -0.02, 0.4)
ROOM = dict(room_lo=(-3.0, -2.5, 0.0), room_hi=(3.0, 2.5, 3.0))


def _tf(k):
    """A slightly different sensor transform per call (row-major 4x4)."""
    tf = np.array(syn.IDENTITY_TF, dtype=np.float32).reshape(4, 4).copy()
    a = np.float32(0.01 * k)
    tf[0, 0], tf[0, 1], tf[1, 0], tf[1, 1] = np.cos(a), -np.sin(a), np.sin(a), np.cos(a)
    tf[0, 3], tf[2, 3] = np.float32(0.02 * k), np.float32(-0.01 * k)
    return tf.reshape(-1)


def _workload(n, p):
    pts, _ = syn.make_scan("vlp16", GT, n_points=p, **ROOM)
    ps = syn.tracking_particles(n, GT, sigma_xy=0.15)
    return ps, pts


@pytest.mark.parametrize("n,p", [(500, 1024), (5000, 4096), (40000, 2048)])
def test_device_update_graph_replays_equal_eager_launches(n, p):
    import torch
    from tsdf_localization_b200.dist import GpuStages

    _, m = common.box_room(small=True)
    ps, pts = _workload(n, p)
    cap = n + n // 8 + 64
    results = {}
    for graphs in (0, 1):
        ev = CudaEvaluator(m)
        ev.tune(capi.TUNE_GRAPHS, graphs)
        st = GpuStages(ev)
        d_pts = torch.from_numpy(pts).cuda()
        d_ps = torch.empty((n, 7), dtype=torch.float32, device="cuda")
        d_out2 = torch.empty((2, cap, 7), dtype=torch.float32, device="cuda")   # double-buffered output, like ShardedSensorUpdate
        d_mean = torch.empty(6, dtype=torch.float32, device="cuda")
        h_ps = torch.from_numpy(ps).cuda()
        runs = []
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            for k in range(8):
                d_ps.copy_(h_ps)
                d_out = d_out2[k & 1]
                d_out.zero_()
                st.update_fused(d_pts, d_ps, n, _tf(k), (0.11 + 0.13 * k) / n, d_out, cap, d_mean)
                n_out, wsum = st.check()
                assert st.last_eval_ms() > 0.0
                runs.append((n_out, wsum, d_ps.cpu().numpy().copy(), d_out[:n_out].cpu().numpy().copy(), d_mean.cpu().numpy().copy()))
        results[graphs] = runs
        captures, replays, note = ev.graph_stats()
        if graphs:
            # call 1 launches kernel by kernel and sizes the scratch buffers (which changes every key), calls 2 and 3 are the
            # first ones with their (output buffer's) key, calls 4 and 5 are recorded, 6 to 8 replay
            assert captures == 2 and replays >= 5 and note == "", (captures, replays, note)
            assert ev.kernel_launches() > 0
        else:
            assert captures == 0 and replays == 0
        ev.close()
    for k, (a, b) in enumerate(zip(results[0], results[1])):
        assert a[0] == b[0] and a[1] == b[1], k
        for x, y in zip(a[2:], b[2:]):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), k
    # the calls do differ from one another (the patched parameters took effect)
    assert not np.array_equal(results[1][1][2], results[1][6][2])
    assert not np.array_equal(results[1][1][3], results[1][6][3])


class _DeviceLoop:
    """tsdfloc_update_device on fixed device buffers (the steady state of a device-resident particle filter)."""

    def __init__(self, m, n, graphs=1, timers=0):
        import torch
        from tsdf_localization_b200.dist import GpuStages
        self.torch, self.n, self.cap = torch, n, n + n // 8 + 64
        self.ev = CudaEvaluator(m)
        self.ev.tune(capi.TUNE_GRAPHS, graphs)
        if timers:
            self.ev.tune(capi.TUNE_STAGE_TIMERS, 1)
        self.st = GpuStages(self.ev)
        self.d_ps = torch.empty((n, 7), dtype=torch.float32, device="cuda")
        self.d_out = torch.empty((self.cap, 7), dtype=torch.float32, device="cuda")
        self.d_mean = torch.empty(6, dtype=torch.float32, device="cuda")
        self.stream = torch.cuda.Stream()

    def update(self, ps, d_pts, tf, u0):
        """Returns the particles with their normalised weights and the resampled set (host copies)."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            self.d_ps.copy_(torch.from_numpy(ps))
            self.st.update_fused(d_pts, self.d_ps, self.n, tf, u0, self.d_out, self.cap, self.d_mean)
            n_out, _ = self.st.check()
            return self.d_ps.cpu().numpy().copy(), self.d_out[:n_out].cpu().numpy().copy()


def test_shape_change_runs_kernel_by_kernel_and_is_recorded_again(oracle):
    """Another scan between replays: that call is launched kernel by kernel, the earlier scan keeps its recording, the new one
    is recorded in turn — every call bit-identical to an evaluator without graphs and, as everywhere, within the north star's
    1e-5 of the oracle's normalised weights, resampled exactly as the oracle resamples those weights."""
    import torch

    _, m = common.box_room(small=True)
    om = common.oracle_map_of(oracle, m)
    n = 400
    ps0, pts_a = _workload(n, 1024)
    _, pts_b = _workload(n, 777)
    d_a, d_b = torch.from_numpy(pts_a).cuda(), torch.from_numpy(pts_b).cuda()
    loop, plain = _DeviceLoop(m, n, graphs=1), _DeviceLoop(m, n, graphs=0)
    u0 = 0.3 / n
    for k, which in enumerate("aaababbbb"):
        pts, d_pts = (pts_a, d_a) if which == "a" else (pts_b, d_b)
        got_ps, got_out = loop.update(ps0, d_pts, syn.IDENTITY_TF, u0)
        want_ps, want_out = plain.update(ps0, d_pts, syn.IDENTITY_TF, u0)
        assert np.array_equal(got_ps.view(np.uint32), want_ps.view(np.uint32)), k
        assert np.array_equal(got_out.view(np.uint32), want_out.view(np.uint32)), k
        ref = oracle.evaluate(om, common.DEFAULT_PARAMS, ps0, pts, syn.IDENTITY_TF)["particles"]
        assert common.rel_err(got_ps[:, 6], ref[:, 6]).max() <= 1e-5, k
        m_ref, parents = oracle.systematic_resample(got_ps[:, 6], u0)
        assert len(got_out) == m_ref and np.array_equal(got_out, got_ps[parents]), k
    captures, replays, note = loop.ev.graph_stats()
    # a (sizes the scratch buffers, which changes every key), a, a recorded; b first call with its key; a replays its
    # recording; b recorded; b, b, b replay
    assert captures == 2 and replays >= 5 and note == "", (captures, replays, note)
    assert plain.ev.graph_stats()[:2] == (0, 0)
    loop.ev.close()
    plain.ev.close()


def test_host_buffer_update_is_launched_kernel_by_kernel():
    """tsdfloc_sensor_update never goes through a graph (its caller waits for the result; measured slower, profiles/r02_graphs.md)."""
    _, m = common.box_room(small=True)
    ev = CudaEvaluator(m)
    ps0, pts = _workload(300, 512)
    before = ev.kernel_launches()
    for _ in range(4):
        ps = ps0.copy()
        ev.evaluate(ps, pts, syn.IDENTITY_TF)
    assert ev.graph_stats()[:2] == (0, 0)
    assert ev.kernel_launches() - before >= 4 * 4
    ev.close()


def test_graphs_are_off_while_stage_timers_run():
    import torch

    _, m = common.box_room(small=True)
    n = 300
    ps0, pts = _workload(n, 512)
    loop = _DeviceLoop(m, n, graphs=1, timers=1)
    d_pts = torch.from_numpy(pts).cuda()
    for _ in range(4):
        loop.update(ps0, d_pts, syn.IDENTITY_TF, 0.5 / n)
    assert loop.ev.graph_stats()[:2] == (0, 0)
    ms = (C.c_float * 4)()
    capi.check(loop.ev._lib, loop.ev.ctx, loop.ev._lib.tsdfloc_stage_times(loop.ev.ctx, ms))
    assert ms[1] > 0.0 and ms[3] > 0.0
    loop.ev.close()
