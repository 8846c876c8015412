"""Dev/profiling driver: a few evaluation launches on a fixed shape (run under ncu, see profiles/README.md)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import common  # noqa: E402
from tsdf_localization_b200 import CudaEvaluator, capi, synthetic as syn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
kind = sys.argv[2] if len(sys.argv) > 2 else "os1-128"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
if kind in ("c4", "c5"):          # the 1.4 GB multi-room map (beyond L2): global localisation / tracking cloud
    m = common.grid_rooms()
    ps, pts, _ = common.config_c4(n) if kind == "c4" else common.config_c5(n)
else:
    spec, m = common.box_room()
    pts, _ = syn.make_scan(kind, syn.GT_POSE)
    ps = syn.tracking_particles(n, syn.GT_POSE)
ev = CudaEvaluator(m)
lib = capi.load_library()
dev = torch.device("cuda:0")
d_ps = torch.from_numpy(ps).to(dev)
d_pts = torch.from_numpy(pts).to(dev)
d_raw = torch.zeros(n, dtype=torch.float32, device=dev)
tf = (C.c_float * 16)(*syn.IDENTITY_TF.tolist())
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
sp = C.c_void_p(ts.cuda_stream)
capi.check(lib, ev.ctx, lib.tsdfloc_set_scan_device(ev.ctx, C.c_void_p(d_pts.data_ptr()), pts.shape[0], sp))
for _ in range(reps):
    capi.check(lib, ev.ctx, lib.tsdfloc_eval_device(ev.ctx, C.c_void_p(d_ps.data_ptr()), n, 0, n, tf, C.c_void_p(d_raw.data_ptr()), sp))
torch.cuda.synchronize()
print("done", float(d_raw.sum()))
