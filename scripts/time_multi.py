"""Times the single-process multi-GPU host-buffer update (tsdfloc_multi_*: what the C++ shim runs with TSDFLOC_DEVICES set)
on the C3 workload for 1, 2, 4, 8 devices. Wall clock around evaluate() + resample_systematic() with host buffers, i.e.
including the H2D of scan + particles to every device and the D2H of weights + resampled particles. One JSON line per N."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import common  # noqa: E402
from tsdf_localization_b200 import CudaEvaluator, MultiGpuEvaluator, synthetic as syn  # noqa: E402


def main():
    _, m = common.box_room()
    ps, pts, _ = common.config_c3()
    n, p = len(ps), len(pts)
    u0 = 0.37 / n
    ref = None
    for world in (1, 2, 4, 8):
        if world > torch.cuda.device_count():
            break
        ev = MultiGpuEvaluator(m, list(range(world))) if world > 1 else CudaEvaluator(m)
        times = []
        for it in range(8):
            mine = ps.copy()
            t0 = time.perf_counter()
            ev.evaluate(mine, pts, syn.IDENTITY_TF)
            out = ev.resample_systematic(u0, capacity=n)
            times.append(1e3 * (time.perf_counter() - t0))
        digest = (mine.tobytes(), out.tobytes())
        if ref is None:
            ref = digest
        print(json.dumps({"devices": world, "particles": n, "points": p, "ms_per_update_host_buffers": float(np.median(times[2:])),
                          "evals_per_s": n * p / (np.median(times[2:]) * 1e-3), "identical_to_1_gpu": digest == ref,
                          "api": "tsdfloc_multi_sensor_update + tsdfloc_multi_resample_systematic" if world > 1 else
                                 "tsdfloc_sensor_update + tsdfloc_resample_systematic"}))
        ev.close()


if __name__ == "__main__":
    main()
