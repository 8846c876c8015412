"""Dev script: what a cold L2 costs the evaluation kernel. CUDA events around tsdfloc_eval_device; one JSON line per
(particles, mode). (Round 2 also tried a linear L2 prefetch of the map ahead of the kernel: no gain, removed —
profiles/r02_cold_l2.md.)

    python scripts/cold_l2.py [out.jsonl] [counts=500,4096,8192,16384,65536]
"""
import ctypes as C
import hashlib
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import common  # noqa: E402
from tsdf_localization_b200 import CudaEvaluator, capi, synthetic as syn  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/cold_l2.jsonl"
    counts = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "500,4096,8192,16384,65536").split(",")]
    _, m = common.box_room()
    ev = CudaEvaluator(m)
    lib = capi.load_library()
    pts, _ = syn.make_scan("os1-128", syn.GT_POSE)
    P = pts.shape[0]
    dev = torch.device("cuda:0")
    d_pts = torch.from_numpy(pts).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tf = (C.c_float * 16)(*syn.IDENTITY_TF.tolist())
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    sp = C.c_void_p(ts.cuda_stream)
    capi.check(lib, ev.ctx, lib.tsdfloc_set_scan_device(ev.ctx, C.c_void_p(d_pts.data_ptr()), P, sp))
    with open(out_path, "w") as f:
        for n in counts:
            ps = syn.tracking_particles(n, syn.GT_POSE)
            d_ps = torch.from_numpy(ps).to(dev)
            d_raw = torch.zeros(n, dtype=torch.float32, device=dev)
            shas = set()
            for mode, do_flush in (("flushed", True), ("resident", False)):
                times = []
                for it in range(8):
                    if do_flush:
                        flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    capi.check(lib, ev.ctx, lib.tsdfloc_eval_device(ev.ctx, C.c_void_p(d_ps.data_ptr()), n, 0, n, tf, C.c_void_p(d_raw.data_ptr()), sp))
                    e1.record()
                    torch.cuda.synchronize()
                    if it >= 2:
                        times.append(e0.elapsed_time(e1))
                shas.add(hashlib.sha256(d_raw.cpu().numpy().tobytes()).hexdigest()[:16])
                row = dict(particles=n, points=P, mode=mode, ms_min=min(times), ms_med=float(np.median(times)))
                f.write(json.dumps(row) + "\n")
                f.flush()
                print(row)
            assert len(shas) == 1


if __name__ == "__main__":
    main()
