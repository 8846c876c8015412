"""Measured ceiling for 4 B gathers on this GPU: prints one JSON line per access shape. The evaluation kernel's
own figure (from ncu: l1tex sectors per request 12.5, lts sectors 6.2e9 per 20.6 ms C3 launch) is compared against these in
DESIGN.md §3."""
import ctypes as C
import json
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import common  # noqa: E402
from tsdf_localization_b200 import CudaEvaluator, capi  # noqa: E402


def main():
    _, m = common.box_room()
    ev = CudaEvaluator(m)
    lib = capi.load_library()
    for label, nbytes in (("32 MB (L2-resident)", 32 << 20), ("77 MB map (L2-resident)", 0)):
        for spread in (0, 32, 16, 13, 8, 4, 1):
            ms, g = C.c_float(0), C.c_uint64(0)
            capi.check(lib, ev.ctx, lib.tsdfloc_probe_gather(ev.ctx, nbytes, spread, 10, C.byref(ms), C.byref(g)))
            sectors = 32 if spread == 0 else spread
            gps = g.value / (ms.value * 1e-3)
            print(json.dumps({"array": label, "lanes_share_sectors": spread or "none (random words)", "gathers_per_s": gps,
                              "sector_GBps": gps / 32 * sectors * 32 / 1e9, "ms": ms.value}))
    ev.close()


if __name__ == "__main__":
    main()
