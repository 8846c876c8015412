"""Host-side checks of the synthetic workload generators (CPU)."""
import numpy as np

import common
from tsdf_localization_b200 import CudaSubVoxelMap, likelihood_init, likelihood_value, synthetic as syn


def test_direct_brick_builder_equals_set_data():
    ext, room, res = (8.0, 6.0, 3.0), 4.0, 0.05
    mk = lambda mn, mx, r, init: CudaSubVoxelMap(*mn, *mx, r, init)   # noqa: E731
    desc, occ, data = syn.grid_rooms_arrays(mk, likelihood_value, likelihood_init(0.1), extent=ext, room=room, resolution=res)
    cells = syn.grid_rooms_cells(likelihood_value, ext, room, res)
    lo, hi = syn.chunk_aligned_bounds((0, 0, 0), ext, res)
    m = CudaSubVoxelMap(*lo, *hi, res, likelihood_init(0.1))
    m.setData(cells)
    assert desc.data_size == m.coef().data_size
    assert np.array_equal(occ, m.rawGridOcc()) and data.tobytes() == m.rawData().tobytes()


def test_reduce_scan_keeps_first_point_per_ring_cell():
    pts, ring = syn.make_scan("vlp16", syn.GT_POSE)
    rp, rr = syn.reduce_scan(pts, ring, 0.256)
    assert 0 < len(rp) < len(pts)
    assert np.all(np.diff(rr) >= 0)                      # ring-major
    keys = set()
    for p, r in zip(rp, rr):
        k = (int(r),) + tuple(np.floor(p / np.float32(0.256)).astype(int))
        assert k not in keys
        keys.add(k)
    # every input (ring, cell) is represented
    all_keys = {(int(r),) + tuple(np.floor(p / np.float32(0.256)).astype(int)) for p, r in zip(pts, ring)}
    assert keys == all_keys


def test_box_room_is_chunk_aligned_and_scan_hits_it():
    spec, m = common.box_room(small=True)
    assert spec.min == (-6.4, -3.2, -3.2) and spec.max == (6.4, 3.2, 6.4)
    assert (m.rawGridOcc() >= 0).sum() * 8000 == m.coef().data_size


def test_committed_bench_lines_are_rank_count_invariant():
    """The round-2 bench lines at 1, 2, 4 and 8 GPUs (profiles/r02_bench_c3*.json, written by `bench.py --gpus N` on the B200
    box) carry sha256 digests of the normalised weight vector and of the resampled particle set, device-resident and through
    the host-buffer C ABI: one update, whatever the number of ranks, produces the same bytes."""
    import json
    from pathlib import Path
    prof = Path(__file__).resolve().parents[1] / "profiles"
    # r02b_*: the same command after the evaluation kernel learned chained scan chunks and the update its CUDA graph — other
    # launches, same bytes
    lines = [(n, json.loads((prof / name).read_text().strip().splitlines()[-1])) for n, name in
             ((1, "r02_bench_c3.json"), (2, "r02_bench_c3_n2.json"), (4, "r02_bench_c3_n4.json"), (8, "r02_bench_c3_n8.json"),
              (1, "r02b_bench_c3.json"), (2, "r02b_bench_c3_n2.json"), (8, "r02b_bench_c3_n8.json"))]
    digests = set()
    for n, d in lines:
        assert d["n_gpus"] == n and d["config"]["particles"] == 65536 and d["config"]["points"] == 131072
        c = d["config"]
        digests.add((c["weights_sha256"], c["resampled_sha256"]))
        digests.add((c["e2e_weights_sha256"], c["e2e_resampled_sha256"]))
    assert len(digests) == 1, digests
