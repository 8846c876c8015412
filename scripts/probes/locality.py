import ctypes as C, sys, json
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import common
from tsdf_localization_b200 import CudaEvaluator, capi, synthetic as syn
_, m = common.box_room()
ev = CudaEvaluator(m, dense_budget_bytes=1)
lib = capi.load_library()
pts, _ = syn.make_scan("os1-128", syn.GT_POSE)
P = pts.shape[0]
dev = torch.device("cuda:0")
d_pts = torch.from_numpy(pts).to(dev)
tf = (C.c_float * 16)(*syn.IDENTITY_TF.tolist())
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); sp = C.c_void_p(ts.cuda_stream)
capi.check(lib, ev.ctx, lib.tsdfloc_set_scan_device(ev.ctx, C.c_void_p(d_pts.data_ptr()), P, sp))
n = 65536
for name, kw in [("tracking sigma .5m/.5rad", {}), ("sigma .05m/.05rad", dict(sigma_xy=0.05, sigma_z=0.01, sigma_rp=0.002, sigma_yaw=0.05)),
                 ("sigma 1mm/1mrad", dict(sigma_xy=0.001, sigma_z=0.001, sigma_rp=0.0001, sigma_yaw=0.001)),
                 ("sorted by yaw", dict(sort=True))]:
    srt = kw.pop("sort", False)
    ps = syn.tracking_particles(n, syn.GT_POSE, **kw)
    if srt:
        ps = ps[np.lexsort((ps[:, 0], ps[:, 5]))]
    d_ps = torch.from_numpy(np.ascontiguousarray(ps)).to(dev)
    d_raw = torch.zeros(n, dtype=torch.float32, device=dev)
    times = []
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        capi.check(lib, ev.ctx, lib.tsdfloc_eval_device(ev.ctx, C.c_void_p(d_ps.data_ptr()), n, 0, n, tf, C.c_void_p(d_raw.data_ptr()), sp))
        e1.record(); torch.cuda.synchronize()
        if it >= 1: times.append(e0.elapsed_time(e1))
    print(json.dumps(dict(case=name, ms=min(times))))
