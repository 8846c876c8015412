"""Host halves of the Wheel / Metropolis / Rejection resamplers (csrc/host_resample.cpp) against the verbatim reference classes
(oracle/_ref) on equally seeded std::mt19937 generators: same parents, wall time of each. CPU only.
    python scripts/time_resamplers_host.py [n ...]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
from oracle_lib import Ref                                                    # noqa: E402
from test_resamplers_host import product_drawn_parents, weighted_cloud        # noqa: E402
from tsdf_localization_b200 import capi                                       # noqa: E402

lib, ref = capi.load_library(), Ref()
for n in [int(a) for a in sys.argv[1:]] or [8192, 65536]:
    ps = weighted_cloud(n, "uniform", 1)
    w = np.ascontiguousarray(ps[:, 6])
    for method, name in ((3, "wheel"), (4, "metropolis(50)"), (5, "rejection")):
        t0 = time.perf_counter()
        _, out, _ = ref.resample_method(method, ps, 7)
        t1 = time.perf_counter()
        d = ref.draws(7, n)
        t2 = time.perf_counter()
        parents = product_drawn_parents(lib, method, w, d, 50)
        t3 = time.perf_counter()
        same = np.array_equal(out[:, 0].astype(np.int64), parents)
        print(f"| {n} | {name} | {(t1 - t0) * 1e3:.2f} | {(t3 - t2) * 1e3:.2f} | {same} |")
