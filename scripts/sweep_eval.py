"""Dev script: evaluation-kernel sweep over particle counts x pairing x register budget (tsdfloc_tune), CUDA events around
tsdfloc_eval_device, L2 flushed before every timed launch. Writes one JSON line per cell; every cell must produce the same
sha256 of the raw weight vector per particle count (all settings are bit-identical by construction).

    python scripts/sweep_eval.py [out.jsonl] [scan=os1-128] [counts=500,2000,8192,16384,65536]
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
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep_eval.jsonl"
    kind = sys.argv[2] if len(sys.argv) > 2 else "os1-128"
    counts = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "500,2000,8192,16384,65536").split(",")]
    reps = 5
    _, m = common.box_room()
    ev = CudaEvaluator(m)
    lib = capi.load_library()
    print("division mode proven:", ev.division_mode())
    pts, _ = syn.make_scan(kind, syn.GT_POSE)
    P = pts.shape[0]
    dev = torch.device("cuda:0")
    d_pts = torch.from_numpy(pts).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tf = (C.c_float * 16)(*syn.IDENTITY_TF.tolist())
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    sp = C.c_void_p(ts.cuda_stream)
    capi.check(lib, ev.ctx, lib.tsdfloc_set_scan_device(ev.ctx, C.c_void_p(d_pts.data_ptr()), P, sp))
    rows = []
    with open(out_path, "w") as f:
        for n in counts:
            ps = syn.tracking_particles(n, syn.GT_POSE)
            d_ps = torch.from_numpy(ps).to(dev)
            d_raw = torch.zeros(n, dtype=torch.float32, device=dev)
            shas = set()
            for pairing in (1, 2):
                for regs in (1, 2):
                    ev.tune(capi.TUNE_EVAL_PAIRING, pairing)
                    ev.tune(capi.TUNE_EVAL_REGISTERS, regs)
                    times = []
                    for it in range(reps + 2):
                        flush.zero_()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        capi.check(lib, ev.ctx, lib.tsdfloc_eval_device(ev.ctx, C.c_void_p(d_ps.data_ptr()), n, 0, n, tf, C.c_void_p(d_raw.data_ptr()), sp))
                        e1.record()
                        torch.cuda.synchronize()
                        if it >= 2:
                            times.append(e0.elapsed_time(e1))
                    sha = hashlib.sha256(d_raw.cpu().numpy().tobytes()).hexdigest()[:16]
                    shas.add(sha)
                    row = dict(particles=n, points=P, pairing={1: "particles", 2: "points"}[pairing], registers={1: 64, 2: 128}[regs],
                               ms_min=min(times), ms_med=float(np.median(times)), gevals_per_s=n * P / min(times) / 1e6, raw_sha=sha)
                    f.write(json.dumps(row) + "\n")
                    f.flush()
                    print(row)
            assert len(shas) == 1, f"settings disagree at n={n}: {shas}"
    print("eval stats", ev.eval_stats())


if __name__ == "__main__":
    main()
