"""world_size-2 (and 3, ragged) CPU tests of the multi-GPU driver's host logic over gloo: slicing, padding, the two
all-gathers, rank-count invariance. The stage kernels are replaced by the CPU oracle (test infrastructure) — the product's
GpuStages is exercised by the -m gpu tests."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


class OracleStages:
    """CPU stand-in for GpuStages with the same call signatures (tensors are CPU float32)."""

    def __init__(self):
        sys.path.insert(0, str(ROOT / "tests"))
        import common
        from oracle_lib import Oracle
        self.common = common
        self.oracle = Oracle()
        _, m = common.box_room(small=True)
        self.omap = common.oracle_map_of(self.oracle, m)
        self.points = None
        self.weights = None
        self.n_out = 0
        self.wsum = 0.0

    def set_scan(self, d_points):
        self.points = d_points.numpy().copy()

    def eval(self, d_particles, n, first, count, tf, d_raw):
        if count == 0:
            return
        ps = d_particles.numpy()[first:first + count]
        out = self.oracle.evaluate(self.omap, self.common.DEFAULT_PARAMS, ps, self.points, np.asarray(tf, dtype=np.float32))
        d_raw[first:first + count] = torch.from_numpy(out["raw"])

    def normalize(self, d_particles, n, d_raw, d_mean):
        raw = d_raw.numpy()[:n]
        self.wsum = float(raw.astype(np.float64).sum())
        w = raw / np.float32(self.wsum) if self.wsum != 0.0 else np.zeros_like(raw)
        d_particles[:n, 6] = torch.from_numpy(w.astype(np.float32))
        self.weights = w.astype(np.float32)
        d_mean[:3] = torch.from_numpy((d_particles[:n, :3].numpy().astype(np.float64) * w[:, None]).sum(0).astype(np.float32))

    def draw(self, d_particles, n, u0, first_out, count_out, d_out):
        m, parents = self.oracle.systematic_resample(self.weights, u0, cap=n + n // 8 + 64)
        self.n_out = m
        slots = np.minimum(np.arange(first_out, first_out + count_out), m - 1)
        d_out[:] = d_particles[torch.from_numpy(parents[slots].astype(np.int64))]

    def check(self):
        if self.wsum == 0.0:
            raise RuntimeError("No particle is valid!")
        return self.n_out, self.wsum


def _workload(n):
    sys.path.insert(0, str(ROOT))
    from tsdf_localization_b200 import synthetic as syn
    gt = (0.4, -0.3, 1.2, 0.01, -0.02, 0.4)
    pts, _ = syn.make_scan("vlp16", gt, room_lo=(-3.0, -2.5, 0.0), room_hi=(3.0, 2.5, 3.0), n_points=256)
    ps = syn.tracking_particles(n, gt, sigma_xy=0.2)
    return ps, pts, syn.CALIB_TF


def _run(rank, world, port, n, u0, ret):
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    from tsdf_localization_b200.dist import ShardedSensorUpdate
    group = None
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    ps, pts, tf = _workload(n)
    upd = ShardedSensorUpdate(OracleStages(), world=world, rank=rank, group=group)
    upd.set_scan(torch.from_numpy(pts))
    d_ps = torch.from_numpy(ps.copy())
    out, mean, n_out, wsum = upd.step(d_ps, n, tf, u0)
    # second update on the resampled set: exercises n != initial n and buffer reuse
    d2 = out.clone()
    out2, _, n_out2, _ = upd.step(d2, n_out, tf, u0)
    ret[rank] = (out.numpy().copy(), d_ps.numpy().copy(), mean.numpy().copy(), n_out, wsum, out2.numpy().copy(), n_out2)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(world, n, u0):
    if world == 1:
        ret = {}
        _run(0, 1, 0, n, u0, ret)
        return ret
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, world, port, n, u0, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    return dict(ret)


def test_shard_covers_everything():
    from tsdf_localization_b200.dist import output_capacity, shard
    for n in (1, 2, 7, 500, 65536, 1_000_000):
        for w in (1, 2, 3, 4, 8):
            seen = 0
            for r in range(w):
                chunk, first, count = shard(n, w, r)
                assert first == min(r * chunk, n) and 0 <= count <= chunk
                assert first == seen or count == 0
                seen += count
            assert seen == n
            cap = output_capacity(n, w)
            assert cap % w == 0 and cap >= n


@pytest.mark.parametrize("world,n", [(2, 500), (3, 97)])
def test_multi_rank_equals_single_rank(world, n):
    u0 = 0.37 / n
    single = _launch(1, n, u0)[0]
    multi = _launch(world, n, u0)
    for r in range(world):
        out, ps, mean, n_out, wsum, out2, n_out2 = multi[r]
        assert n_out == single[3] and n_out2 == single[6]
        assert wsum == single[4]
        assert out.tobytes() == single[0].tobytes(), f"rank {r}: resampled particles differ from the 1-rank run"
        assert ps.tobytes() == single[1].tobytes(), f"rank {r}: normalised weights differ"
        assert out2.tobytes() == single[5].tobytes()
        np.testing.assert_array_equal(mean, single[2])
