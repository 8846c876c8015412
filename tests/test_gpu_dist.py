"""GPU tests of the particle-sharded update (north star item 4): R-rank results must be IDENTICAL to the 1-rank result.

* On one GPU the ranks are emulated by threads, each with its own tsdfloc ctx + stream on cuda:0, and an all-gather that
  copies between the ranks' buffers at a thread barrier — the product's GpuStages / ShardedSensorUpdate code is what runs.
* With >= 2 GPUs (gpurun --gpus 2) a real torchrun/NCCL run of the same check is launched as a subprocess.
"""
import os
import subprocess
import sys
import threading
from pathlib import Path

import numpy as np
import pytest

import common
from tsdf_localization_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _single(m, ps, pts, tf, u0):
    import torch
    from tsdf_localization_b200 import CudaEvaluator
    from tsdf_localization_b200.dist import GpuStages, ShardedSensorUpdate
    dev = torch.device("cuda", 0)
    ev = CudaEvaluator(m)
    upd = ShardedSensorUpdate(GpuStages(ev), device=dev)
    upd.set_scan(torch.from_numpy(pts).to(dev))
    d = torch.from_numpy(ps).to(dev)
    out, mean, n_out, wsum = upd.step(d, len(ps), tf, u0)
    res = (out.cpu().numpy().copy(), d.cpu().numpy().copy(), mean.cpu().numpy().copy(), n_out, wsum)
    ev.close()
    return res


@pytest.mark.parametrize("world,n", [(2, 4096), (3, 1000), (8, 5000)])
def test_thread_emulated_ranks_equal_single_rank(world, n):
    import torch
    from tsdf_localization_b200 import CudaEvaluator
    from tsdf_localization_b200.dist import GpuStages, ShardedSensorUpdate

    _, m = common.box_room()
    pts, _ = syn.make_scan("vlp16", syn.GT_POSE, n_points=4000)
    ps = syn.tracking_particles(n, syn.GT_POSE)
    tf, u0 = syn.CALIB_TF, 0.37 / n
    single = _single(m, ps, pts, tf, u0)

    dev = torch.device("cuda", 0)
    bar = threading.Barrier(world)
    pending = {}
    results, errors = {}, []

    def make_all_gather(rank):
        def all_gather(out, inp):
            # publish my slice, wait for everyone, copy the peers' slices into my output buffer
            torch.cuda.current_stream().synchronize()
            pending[rank] = inp
            bar.wait()
            k = inp.numel()
            for r in range(world):
                if r != rank:
                    out[r * k:(r + 1) * k].copy_(pending[r])
            torch.cuda.current_stream().synchronize()
            bar.wait()
        return all_gather

    def run(rank):
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(torch.cuda.Stream(device=dev)):
                ev = CudaEvaluator(m)
                upd = ShardedSensorUpdate(GpuStages(ev), world=world, rank=rank, device=dev, all_gather=make_all_gather(rank))
                upd.set_scan(torch.from_numpy(pts).to(dev))
                d = torch.from_numpy(ps).to(dev)
                out, mean, n_out, wsum = upd.step(d, n, tf, u0)
                results[rank] = (out.cpu().numpy().copy(), d.cpu().numpy().copy(), mean.cpu().numpy().copy(), n_out, wsum)
                ev.close()
        except Exception as e:  # noqa: BLE001
            errors.append((rank, repr(e)))
            bar.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(300)
    assert not errors, errors
    for r in range(world):
        out, d, mean, n_out, wsum = results[r]
        assert n_out == single[3] and wsum == single[4]
        assert d.tobytes() == single[1].tobytes(), f"rank {r}: normalised weights differ from the 1-rank run"
        assert out.tobytes() == single[0].tobytes(), f"rank {r}: resampled particles differ from the 1-rank run"
        assert mean.tobytes() == single[2].tobytes()


def test_peer_store_kernels_on_one_gpu():
    """tsdfloc_eval_device_peers / tsdfloc_draw_device_peers with the "peers" being two buffers on the same GPU: every rank's
    kernel stores its slice into both weight vectors / particle buffers, so after both ranks ran, both copies hold the full
    result — no all-gather anywhere."""
    import torch
    from tsdf_localization_b200 import CudaEvaluator
    from tsdf_localization_b200.dist import GpuStages, output_capacity, shard

    _, m = common.box_room()
    n, world = 5001, 2
    pts, _ = syn.make_scan("vlp16", syn.GT_POSE, n_points=4000)
    ps = syn.tracking_particles(n, syn.GT_POSE)
    tf, u0 = syn.CALIB_TF, 0.37 / n
    single = _single(m, ps, pts, tf, u0)
    dev = torch.device("cuda", 0)
    evs = [CudaEvaluator(m) for _ in range(world)]
    st = [GpuStages(e) for e in evs]
    chunk = shard(n, world, 0)[0]
    ocap = output_capacity(n, world)
    ochunk = ocap // world
    raws = [torch.zeros(world * chunk, device=dev) for _ in range(world)]
    outs = [torch.zeros((ocap, 7), device=dev) for _ in range(world)]
    parts = [torch.from_numpy(ps).to(dev) for _ in range(world)]
    means = [torch.zeros(8, device=dev) for _ in range(world)]
    d_pts = torch.from_numpy(pts).to(dev)
    for r in range(world):
        st[r].set_scan(d_pts)
        _, first, count = shard(n, world, r)
        st[r].eval_peers(parts[r], n, first, count, tf, raws[r], [t.data_ptr() for t in raws])
    torch.cuda.synchronize()
    assert torch.equal(raws[0], raws[1])
    for r in range(world):
        st[r].normalize(parts[r], n, raws[r], means[r])
        slot = r * ochunk * 28
        st[r].draw_peers(parts[r], n, u0, r * ochunk, ochunk, outs[r][r * ochunk:(r + 1) * ochunk], [t.data_ptr() + slot for t in outs])
    torch.cuda.synchronize()
    for r in range(world):
        n_out, wsum = st[r].check()
        assert n_out == single[3] and wsum == single[4]
        assert parts[r].cpu().numpy().tobytes() == single[1].tobytes()
        assert outs[r][:n_out].cpu().numpy().tobytes() == single[0].tobytes()
    for e in evs:
        e.close()


def test_nccl_two_ranks_equal_single_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", str(ROOT / "tests" / "dist_nccl_check.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "NCCL_CHECK_OK" in res.stdout


@pytest.mark.parametrize("n", [4096, 5001])
def test_single_process_multi_gpu_equals_single_gpu(n):
    """tsdfloc_multi_* (one process, one ctx per device, peer stores + CUDA events): bit-identical to the single-GPU calls.
    Runs over the real devices when there are >= 2, and always over the same device used three times (three ranks)."""
    import torch
    from tsdf_localization_b200 import CudaEvaluator, MultiGpuEvaluator
    _, m = common.box_room()
    pts, _ = syn.make_scan("vlp16", syn.GT_POSE, n_points=5000)
    ps = syn.tracking_particles(n, syn.GT_POSE)
    tf, u0 = syn.CALIB_TF, 0.37 / n
    ev = CudaEvaluator(m)
    want = ps.copy()
    pose1 = ev.evaluate(want, pts, tf)
    out1 = ev.resample_systematic(u0, capacity=n + n // 8 + 64)
    ev.close()
    layouts = [[0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        layouts.append(list(range(min(torch.cuda.device_count(), 8))))
    for devices in layouts:
        mev = MultiGpuEvaluator(m, devices)
        for _ in range(2):                      # twice: buffers and peer tables are reused
            got = ps.copy()
            pose = mev.evaluate(got, pts, tf)
            out = mev.resample_systematic(u0, capacity=n + n // 8 + 64)
            assert got.tobytes() == want.tobytes(), f"normalised weights differ on devices {devices}"
            assert out.tobytes() == out1.tobytes(), f"resampled particles differ on devices {devices}"
            assert pose.position == pose1.position and pose.rpy == pose1.rpy
        mev.close()
    # the "No particle is valid!" condition (all weights 0) is detected on every device and reported like the 1-GPU call
    mev = MultiGpuEvaluator(m, [0, 0], a_range=0.0)
    far = ps.copy()
    far[:, :3] += 500.0
    with pytest.raises(RuntimeError, match="No particle is valid!"):
        mev.evaluate(far, pts, tf)
    mev.close()


def test_single_process_multi_gpu_cloud_equals_single_gpu():
    """tsdfloc_multi_sensor_update_cloud (TSDFEvaluator::evaluateParticles over several devices): reduction on the first device,
    reduced scan to the others, sharded update — bit-identical to tsdfloc_sensor_update_cloud on one GPU."""
    import torch
    from tsdf_localization_b200 import CudaEvaluator, MultiGpuEvaluator
    from test_reduce_oracle import scan_with_rings
    _, m = common.box_room()
    pts, ring = scan_with_rings("vlp16", near=200, shuffle=True)
    n = 3001
    ps = syn.tracking_particles(n, syn.GT_POSE)
    tf, u0 = syn.CALIB_TF, 0.41 / n
    ev = CudaEvaluator(m)
    want = ps.copy()
    pose1, used1 = ev.evaluate_cloud(want, pts, ring, tf, n_rings=16)
    out1 = ev.resample_systematic(u0, capacity=n + n // 8 + 64)
    ev.close()
    assert used1 > 0
    layouts = [[0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        layouts.append(list(range(min(torch.cuda.device_count(), 8))))
    for devices in layouts:
        mev = MultiGpuEvaluator(m, devices)
        for _ in range(2):
            got = ps.copy()
            pose, used = mev.evaluate_cloud(got, pts, ring, tf, n_rings=16)
            out = mev.resample_systematic(u0, capacity=n + n // 8 + 64)
            assert used == used1
            assert got.tobytes() == want.tobytes(), f"normalised weights differ on devices {devices}"
            assert out.tobytes() == out1.tobytes(), f"resampled particles differ on devices {devices}"
            assert pose.position == pose1.position and pose.rpy == pose1.rpy
        # a cloud whose every point is nearer than 1 m reduces to nothing: weights untouched, default pose
        near = (pts[:50] * 0.0 + 0.1).astype(np.float32)
        got = ps.copy()
        pose, used = mev.evaluate_cloud(got, near, ring[:50], tf, n_rings=16)
        assert used == 0 and got.tobytes() == ps.tobytes()
        mev.close()
