"""Dev script: time the device-resident stages on a config (not the bench contract; see bench.py)."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import common  # noqa: E402
from tsdf_localization_b200 import CudaEvaluator, capi, synthetic as syn  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    kind = sys.argv[2] if len(sys.argv) > 2 else "os1-128"
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    import os
    res = float(os.environ.get("QUICK_RES", "0.05"))       # QUICK_RES=0.064: the mapping pipeline's own resolution (sub_dim 16)
    spec, m = common.box_room(resolution=res)
    ev = CudaEvaluator(m)
    lib = capi.load_library()
    pts, _ = syn.make_scan(kind, syn.GT_POSE)
    ps = syn.tracking_particles(n, syn.GT_POSE)
    dev = torch.device("cuda:0")
    d_ps = torch.from_numpy(ps).to(dev)
    d_pts = torch.from_numpy(pts).to(dev)
    d_raw = torch.zeros(n, dtype=torch.float32, device=dev)
    d_out = torch.zeros((n + n // 8 + 64, 7), dtype=torch.float32, device=dev)
    d_mean = torch.zeros(8, dtype=torch.float32, device=dev)
    tf = (C.c_float * 16)(*syn.IDENTITY_TF.tolist())
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    sp = C.c_void_p(ts.cuda_stream)   # NULL would mean "the ctx's own stream"
    capi.check(lib, ev.ctx, lib.tsdfloc_set_scan_device(ev.ctx, C.c_void_p(d_pts.data_ptr()), pts.shape[0], sp))
    P = pts.shape[0]
    for it in range(reps + 2):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        capi.check(lib, ev.ctx, lib.tsdfloc_eval_device(ev.ctx, C.c_void_p(d_ps.data_ptr()), n, 0, n, tf, C.c_void_p(d_raw.data_ptr()), sp))
        e[1].record()
        capi.check(lib, ev.ctx, lib.tsdfloc_normalize_device(ev.ctx, C.c_void_p(d_ps.data_ptr()), n, C.c_void_p(d_raw.data_ptr()), C.c_void_p(d_mean.data_ptr()), sp))
        e[2].record()
        capi.check(lib, ev.ctx, lib.tsdfloc_draw_device(ev.ctx, C.c_void_p(d_ps.data_ptr()), n, 0.37 / n, 0, d_out.shape[0], C.c_void_p(d_out.data_ptr()), None, sp))
        e[3].record()
        torch.cuda.synchronize()
        n_out = C.c_uint64()
        ws = C.c_double()
        capi.check(lib, ev.ctx, lib.tsdfloc_check(ev.ctx, C.byref(n_out), C.byref(ws), sp))
        t_eval, t_norm, t_draw = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])
        if it == reps + 1:
            st = (C.c_uint64 * 4)()
            lib.tsdfloc_eval_stats(ev.ctx, st)
            print(f"eval blocks {st[0]} folded {st[1]} ({100.0 * st[1] / max(st[0], 1):.2f}%) tie-folds {st[2]}")
        if it == reps + 1:
            import hashlib
            print("raw sha", hashlib.sha256(d_raw.cpu().numpy().tobytes()).hexdigest()[:16], "variant",
                  {k: v for k, v in __import__("os").environ.items() if k.startswith("TSDFLOC_")})
        print(f"it{it}: N={n} P={P} eval {t_eval:.3f} ms ({n * P / t_eval / 1e6:.1f} Geval/s)  normalise {t_norm:.3f} ms  draw {t_draw:.3f} ms  n_out={n_out.value} sum={ws.value:.6g}")


if __name__ == "__main__":
    main()
