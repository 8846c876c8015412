"""Dev script: per-stage device times (tsdfloc_stage_times) and host wall-clock of the host-buffer update for c1 / c2 / c3."""
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import bench  # noqa: E402
from tsdf_localization_b200 import CudaEvaluator, capi  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/stage_times.jsonl"
    lib = capi.load_library()
    with open(out_path, "w") as f:
        for name in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("c1", "c2", "c3")):
            ps, pts, tf = bench.workload_inputs(name)
            m = bench.product_map(name)
            ev = CudaEvaluator(m)
            n, p = len(ps), len(pts)
            h_ps = torch.from_numpy(ps).pin_memory()
            h_pts = torch.from_numpy(pts).pin_memory()
            cap = n + n // 8 + 64
            h_out = torch.empty((cap, 7), dtype=torch.float32).pin_memory()
            tfc = (C.c_float * 16)(*[float(v) for v in tf])
            mean = (C.c_float * 6)()
            n_out = C.c_uint64(0)
            u0 = 0.37 / n

            def update():
                capi.check(lib, ev.ctx, lib.tsdfloc_sensor_update(ev.ctx, C.c_void_p(h_ps.data_ptr()), n, C.c_void_p(h_pts.data_ptr()), p, tfc, mean))
                t1 = time.perf_counter()
                capi.check(lib, ev.ctx, lib.tsdfloc_resample_systematic(ev.ctx, C.c_float(u0), C.c_void_p(h_out.data_ptr()), cap, C.byref(n_out), None))
                return t1

            for timers in (0, 1):
                ev.tune(capi.TUNE_STAGE_TIMERS, timers)
                walls, splits = [], []
                for it in range(12):
                    t0 = time.perf_counter()
                    t1 = update()
                    t2 = time.perf_counter()
                    if it >= 2:
                        walls.append(1e3 * (t2 - t0))
                        splits.append((1e3 * (t1 - t0), 1e3 * (t2 - t1)))
                row = dict(workload=name, particles=n, points=p, stage_timers=bool(timers), wall_ms_best=min(walls), wall_ms_median=float(np.median(walls)),
                           sensor_update_ms=min(s[0] for s in splits), resample_ms=min(s[1] for s in splits))
                if timers:
                    ms = (C.c_float * 4)()
                    capi.check(lib, ev.ctx, lib.tsdfloc_stage_times(ev.ctx, ms))
                    row.update(init_kernel_ms=ms[0], exec_kernel_ms=ms[1], weight_update_ms=ms[2], resampling_ms=ms[3])
                print(row)
                f.write(json.dumps(row) + "\n")
            ev.close()


if __name__ == "__main__":
    main()
