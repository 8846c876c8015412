"""torchrun worker for tests/test_gpu_dist.py::test_nccl_two_ranks_equal_single_rank (needs >= 2 GPUs).

Every rank runs the NCCL-sharded update and, on its own GPU, the unsharded one; the results must be byte-identical
(normalised weights, resampled particles, mean pose, n_out), for two consecutive updates.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import common  # noqa: E402
from tsdf_localization_b200 import CudaEvaluator, synthetic as syn  # noqa: E402
from tsdf_localization_b200.dist import GpuStages, ShardedSensorUpdate  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    _, m = common.box_room()
    pts, _ = syn.make_scan("vlp16", syn.GT_POSE, n_points=6000)
    for n in (4096, 5001):
        ps = syn.tracking_particles(n, syn.GT_POSE)
        tf, u0 = syn.CALIB_TF, 0.37 / n
        # both transports: NCCL all-gathers, and the fused kernels storing straight into the peers' symmetric-memory buffers
        for fused in (False, True):
            ev1, evw = CudaEvaluator(m, device=local), CudaEvaluator(m, device=local)
            one = ShardedSensorUpdate(GpuStages(ev1), device=dev)
            many = ShardedSensorUpdate(GpuStages(evw), world=world, rank=rank, device=dev, max_particles=n, fused=fused)
            assert many.transport == ("fused_p2p" if fused else "all_gather"), many.transport
            d_pts = torch.from_numpy(pts).to(dev)
            one.set_scan(d_pts)
            many.set_scan(d_pts)
            a, b = torch.from_numpy(ps).to(dev), torch.from_numpy(ps).to(dev)
            for it in range(3):
                o1, m1, n1, w1 = one.step(a, len(a), tf, u0)
                o2, m2, n2, w2 = many.step(b, len(b), tf, u0)
                assert n1 == n2 and w1 == w2, (fused, n1, n2, w1, w2)
                assert torch.equal(a, b), f"normalised weights differ (fused={fused})"
                assert torch.equal(o1, o2), f"resampled particles differ (fused={fused})"
                assert torch.equal(m1, m2), f"mean pose differs (fused={fused})"
                a, b = o1.clone(), o2.clone()
            torch.cuda.synchronize(dev)
            dist.barrier()
            ev1.close()
            evw.close()
    dist.barrier()
    if rank == 0:
        print("NCCL_CHECK_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
