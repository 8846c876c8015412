"""Dev/profile script: time the GPU scan reduction (device-resident and host-buffer) against the reference's own
evaluateParticles reduction on the host CPU (oracle/_ref, verbatim). Prints one JSON line per cloud."""
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import common  # noqa: E402
from oracle_lib import Oracle, Ref, ref_available  # noqa: E402
from tsdf_localization_b200 import CudaEvaluator, capi, synthetic as syn  # noqa: E402


def main():
    _, m = common.box_room(small=True)
    ev = CudaEvaluator(m)
    lib = capi.load_library()
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream()
    for kind, n_rings, cell in (("vlp16", 16, 0.064), ("os1-128", 128, 0.064), ("os1-128", 128, 0.256)):
        pts, ring = syn.make_scan(kind, syn.GT_POSE)
        ring = np.ascontiguousarray(ring, dtype=np.int32)
        n = len(pts)
        d_pts = torch.from_numpy(pts).to(dev)
        d_ring = torch.from_numpy(ring).to(dev)
        d_out = torch.empty((n, 3), dtype=torch.float32, device=dev)
        d_src = torch.empty(n, dtype=torch.int32, device=dev)
        sp = C.c_void_p(stream.cuda_stream)
        n_out = C.c_uint64(0)
        l0 = ev.kernel_launches()

        def dev_run():
            capi.check(lib, ev.ctx, lib.tsdfloc_reduce_scan_device(ev.ctx, C.c_void_p(d_pts.data_ptr()), C.c_void_p(d_ring.data_ptr()), n,
                                                                  C.c_double(cell), n_rings, 0, C.c_void_p(d_out.data_ptr()),
                                                                  C.c_void_p(d_src.data_ptr()), sp))
        with torch.cuda.stream(stream):
            for _ in range(5):
                dev_run()
            capi.check(lib, ev.ctx, lib.tsdfloc_reduce_result(ev.ctx, C.byref(n_out), sp))
            launches = (ev.kernel_launches() - l0) // 5
            ts = []
            for _ in range(20):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                dev_run()
                e1.record(stream)
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
        t_dev = float(np.median(ts))
        th = []
        for _ in range(10):
            t0 = time.perf_counter()
            out = ev.reduce_scan(pts, ring, cell, n_rings=n_rings)
            th.append(1e3 * (time.perf_counter() - t0))
        want, _ = Oracle().reduce_scan(pts, ring, cell, n_rings=n_rings)
        assert out.tobytes() == want.tobytes() and int(n_out.value) == len(want)
        t_ref = None
        if ref_available() and n_rings <= 64:
            r = Ref()
            r.reduce_scan(pts, ring, cell)
            tr = []
            for _ in range(5):
                t0 = time.perf_counter()
                r.reduce_scan(pts, ring, cell)
                tr.append(1e3 * (time.perf_counter() - t0))
            t_ref = float(np.median(tr))
        elif ref_available():
            # the reference indexes 64 ring buckets: fold the rings so that it can run at all (timing only)
            r = Ref()
            tr = []
            for _ in range(5):
                t0 = time.perf_counter()
                r.reduce_scan(pts, ring % 64, cell)
                tr.append(1e3 * (time.perf_counter() - t0))
            t_ref = float(np.median(tr))
        print(json.dumps({"cloud": kind, "points": n, "cell": cell, "reduced": len(want), "gpu_device_ms": t_dev, "gpu_launches": int(launches),
                          "gpu_host_buffers_ms": float(np.median(th)), "reference_cpu_ms": t_ref,
                          "note": "reference = TSDFEvaluator::evaluateParticles' reduction, verbatim, 1 thread (it is serial); includes its cloud copy"}))
    ev.close()


if __name__ == "__main__":
    main()
