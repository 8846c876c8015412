"""Dev script: cost of the exact-order CDF redo (K3) — systematic resampling of a weighted cloud whose fp64 running sum rounds
(weights spanning 80 binades) against a cloud of the same size whose sum is exact. Wall-clock of tsdfloc_resample_particles from
pinned host buffers, best of 5; the difference is the redo. Also checks the parents against the oracle's serial loop."""
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import common  # noqa: E402
from oracle_lib import Oracle  # noqa: E402
from tsdf_localization_b200 import CudaEvaluator, capi  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/fallback_timing.jsonl"
    _, m = common.box_room(small=True)
    ev = CudaEvaluator(m)
    lib = capi.load_library()
    oracle = Oracle()
    rng = np.random.default_rng(3)
    with open(out_path, "w") as f:
        for n in (1 << 16, 1 << 20):
            rows = {}
            for kind in ("narrow", "wide"):
                w = rng.random(n) + 0.5 if kind == "narrow" else 10.0 ** rng.uniform(-24.0, 0.0, n)
                w = (w / w.sum()).astype(np.float32)
                ps = torch.zeros((n, 7), dtype=torch.float32).pin_memory()
                ps[:, 0] = torch.arange(n, dtype=torch.float32)
                ps[:, 6] = torch.from_numpy(w)
                cap = n + n // 8 + 64
                out = torch.empty((cap, 7), dtype=torch.float32).pin_memory()
                par = torch.empty(cap, dtype=torch.int32).pin_memory()
                n_out = C.c_uint64(0)
                u0 = 0.37 / n
                times = []
                for _ in range(6):
                    t0 = time.perf_counter()
                    capi.check(lib, ev.ctx, lib.tsdfloc_resample_particles(ev.ctx, C.c_void_p(ps.data_ptr()), n, C.c_float(u0),
                                                                           C.c_void_p(out.data_ptr()), cap, C.byref(n_out), C.c_void_p(par.data_ptr())))
                    times.append(1e3 * (time.perf_counter() - t0))
                m_ref, parents_ref = oracle.systematic_resample(w, u0, cap=cap)
                ok = (n_out.value == m_ref) and np.array_equal(par.numpy()[:m_ref].astype(np.uint32), parents_ref)
                exact = bool(lib.tsdfloc_last_cdf_was_exact(ev.ctx))
                rows[kind] = dict(n=n, weights=kind, ms_best=min(times[1:]), parents_match_oracle=bool(ok), parallel_scan_was_exact=exact)
                print(rows[kind])
                f.write(json.dumps(rows[kind]) + "\n")
            print(f"n={n}: redo costs {rows['wide']['ms_best'] - rows['narrow']['ms_best']:.3f} ms")


if __name__ == "__main__":
    main()
