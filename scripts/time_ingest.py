"""Times the map ingest (mapping-pipeline chunks -> two-level sparse map): host loop (tsdfloc_map_from_chunks, the reference's
createTSDFMap structure) against the GPU ingest (tsdfloc_map_from_chunks_gpu), and checks that the arrays are identical."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_map_ingest import synthetic_chunks  # noqa: E402
from tsdf_localization_b200 import CudaSubVoxelMap  # noqa: E402


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    pos = [(x, y, z) for x in range(side) for y in range(side) for z in range(2)]
    centre = (side * 32 * 64.0, side * 32 * 64.0, 64 * 64.0)
    data = synthetic_chunks(pos, seed=1, radius_mm=side * 20 * 64.0, centre_mm=centre)
    t0 = time.perf_counter()
    host = CudaSubVoxelMap.from_chunks(pos, data, 0.1)
    t_host = time.perf_counter() - t0
    CudaSubVoxelMap.from_chunks(pos[:1], data[:1], 0.1, device=0)          # context creation outside the timing
    t0 = time.perf_counter()
    gpu = CudaSubVoxelMap.from_chunks(pos, data, 0.1, device=0)
    t_gpu = time.perf_counter() - t0
    same = (np.array_equal(host.rawGridOcc(), gpu.rawGridOcc()) and host.rawData().tobytes() == gpu.rawData().tobytes()
            and host.free_map().tobytes() == gpu.free_map().tobytes())
    print(json.dumps({"chunks": len(pos), "raw_mb": data.nbytes / 1e6, "map_mb": host.dataBytes() / 1e6, "free_points": len(host.free_map()),
                      "host_ingest_s": t_host, "gpu_ingest_s": t_gpu, "identical": bool(same)}))


if __name__ == "__main__":
    main()
