"""Dev/profile script: the six resamplers of mcl_3d's resampling_method switch on a B200, host cloud in -> resampled host cloud
out (what Resampler::resample(ParticleCloud&) does), against the verbatim reference classes on the host CPU (oracle/_ref) with
equal seeds. Prints one JSON line per method and particle count; `same` = identical output particles.
    python scripts/time_resamplers_gpu.py [n ...]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import common  # noqa: E402
from oracle_lib import Ref  # noqa: E402
from test_resamplers_host import product_drawn_parents, weighted_cloud  # noqa: E402
from tsdf_localization_b200 import (CudaEvaluator, MetropolisResampler, RejectionResampler, ResidualResampler,  # noqa: E402
                                    ResidualSystematicResampler, SystematicResampler, WheelResampler, capi)


def best(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), out


def main():
    ref = Ref()
    _, m = common.box_room(small=True)
    ev = CudaEvaluator(m)
    for n in [int(a) for a in sys.argv[1:]] or [8192, 65536]:
        ps = weighted_cloud(n, "uniform", 1)
        seed = 7
        rows = []
        # Systematic / ResidualSystematic: the reference's draw is handed over; Residual: its index stream
        _, out_ref, u0 = ref.systematic_resample(ps, seed)
        t_ref, _ = best(lambda: ref.systematic_resample(ps, seed), 3)
        t_gpu, out = best(lambda: SystematicResampler(ev).resample(ps, u0=u0), 5)
        rows.append(("systematic", t_ref, t_gpu, np.array_equal(out, out_ref)))
        _, out_ref, u = ref.resample_method(2, ps, seed)
        t_ref, _ = best(lambda: ref.resample_method(2, ps, seed), 3)
        t_gpu, out = best(lambda: ResidualSystematicResampler(ev).resample(ps, u0=u), 5)
        rows.append(("residual_systematic", t_ref, t_gpu, np.array_equal(out, out_ref)))
        _, out_ref, _ = ref.resample_method(1, ps, seed)
        t_ref, _ = best(lambda: ref.resample_method(1, ps, seed), 3)
        draws = ref.uniform_index_draws(seed, n, 64 * n + 1024)
        t_gpu, out = best(lambda: ResidualResampler(ev).resample(ps, index_draws=draws), 1)   # Python callback per draw: not a timing
        rows.append(("residual (python draw callbacks)", t_ref, t_gpu, np.array_equal(out, out_ref)))
        for method, cls, name in ((3, WheelResampler, "wheel"), (4, MetropolisResampler, "metropolis(50)"), (5, RejectionResampler, "rejection")):
            if method == 3 and n > 16384:
                # the reference's O(n^2) wheel takes 19 s at 65,536 (profiles/r02c_resamplers_host.md): not on GPU-box time; the
                # expected output comes from the host half, which tests/test_resamplers_host.py pins to the reference class
                parents = product_drawn_parents(capi.load_library(), 3, np.ascontiguousarray(ps[:, 6]), ref.draws(seed, n))
                t_ref, out_ref = float("nan"), ps[parents]
            else:
                t_ref, (_, out_ref, _) = best(lambda: ref.resample_method(method, ps, seed), 1 if method == 3 else 3)
            rs = cls(ev)

            def run():
                d = ref.draws(seed, n)          # std::mt19937(seed) + the reference's distribution objects, native callbacks
                return rs.resample(ps, draws=d.source())
            t_gpu, out = best(run, 5)
            rows.append((name, t_ref, t_gpu, np.array_equal(out, out_ref)))
        for name, t_ref, t_gpu, same in rows:
            print(json.dumps({"particles": n, "method": name, "reference_cpu_ms": None if t_ref != t_ref else round(t_ref, 3), "b200_host_to_host_ms": round(t_gpu, 3),
                              "same": bool(same)}), flush=True)
    ev.close()


if __name__ == "__main__":
    main()
