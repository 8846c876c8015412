"""Map ingest: tsdfloc_map_from_chunks against the reference's own createTSDFMap
(include/tsdf_localization/map/map_util.h:17-154), compiled verbatim over an in-memory stand-in for the HDF5 file
(oracle/ref_stubs/highfive). Bit-exact geometry, brick table, voxel payload and free-space points. CPU only."""
import numpy as np
import pytest

from oracle_lib import Ref, ref_available
from tsdf_localization_b200 import CudaSubVoxelMap

needs_ref = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")


def synthetic_chunks(chunk_pos, seed=0, radius_mm=2500.0, centre_mm=(1000.0, 2000.0, 500.0)):
    """64^3 chunks of packed {int16 tsdf_mm, int16 weight}: signed distance to a sphere, clipped like the mapping pipeline
    (+-600 mm), with untouched (weight 0) voxels and a band of free-space voxels (|tsdf| == 600) around it."""
    rng = np.random.default_rng(seed)
    out = np.zeros((len(chunk_pos), 64, 64, 64), dtype=np.uint32)
    ax = np.arange(64)
    for c, (cx, cy, cz) in enumerate(chunk_pos):
        x = ((64 * cx + ax) * 64.0)[:, None, None]
        y = ((64 * cy + ax) * 64.0)[None, :, None]
        z = ((64 * cz + ax) * 64.0)[None, None, :]
        d = np.sqrt((x - centre_mm[0]) ** 2 + (y - centre_mm[1]) ** 2 + (z - centre_mm[2]) ** 2) - radius_mm
        value = np.clip(np.rint(d), -600, 600).astype(np.int16)
        weight = np.where(np.abs(d) < 1500, rng.integers(1, 20, size=d.shape), 0).astype(np.int16)
        weight[rng.random(d.shape) < 0.1] = 0
        out[c] = (value.view(np.uint16).astype(np.uint32)) | (weight.view(np.uint16).astype(np.uint32) << 16)
    return out


def assert_same_map(ref, h, m):
    coef, occ, data = ref.map_arrays(h)
    mine = m.coef()
    for f in ("resolution", "init_value", "up_dim_2", "sub_dim", "sub_dim_2", "grid_occ_size", "data_size"):
        assert getattr(mine, f) == getattr(coef, f), f
    for f in ("dim", "min", "max", "up_dim"):
        assert list(getattr(mine, f)) == list(getattr(coef, f)), f
    assert np.array_equal(m.rawGridOcc(), occ)
    assert m.rawData().tobytes() == data.tobytes()


@needs_ref
@pytest.mark.parametrize("chunk_pos", [
    [(0, 0, 0)],
    [(0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 1, 0)],
    [(-1, 0, 0), (0, 0, -1), (0, 0, 0), (-1, -1, -1), (1, 0, 0)],     # negative chunk coordinates, name order != list order
    [(2, 1, 0), (10, 1, 0)],                                           # "10_1_0" sorts before "2_1_0"; box still starts at 0
], ids=["one", "four", "negative", "sparse"])
def test_map_from_chunks_matches_createTSDFMap(chunk_pos, ):
    ref = Ref()
    centre = (1000.0, 2000.0, 500.0) if chunk_pos[0] != (2, 1, 0) else (9000.0, 6000.0, 2000.0)
    data = synthetic_chunks(chunk_pos, seed=len(chunk_pos), centre_mm=centre)
    h, free_ref = ref.create_tsdf_map(chunk_pos, data, 0.1)
    m = CudaSubVoxelMap.from_chunks(chunk_pos, data, 0.1)
    assert_same_map(ref, h, m)
    assert m.coef().data_size > 0 and len(free_ref) > 0
    assert m.free_map().tobytes() == free_ref.tobytes()
    # lookups through the reference's own getEntry agree with the arrays we built (spot check incl. unmapped space)
    rng = np.random.default_rng(1)
    lo, hi = np.array(list(m.coef().min)), np.array(list(m.coef().max))
    q = rng.uniform(lo, hi, size=(20000, 3)).astype(np.float32)
    assert np.count_nonzero(ref.get_entries(h, q)) > 0
    ref.map_destroy(h)


@needs_ref
def test_other_sigma_and_empty_chunk():
    ref = Ref()
    pos = [(0, 0, 0), (0, 0, 1)]
    data = synthetic_chunks(pos, seed=3)
    data[1] = 0                                   # a chunk without a single touched voxel
    h, free_ref = ref.create_tsdf_map(pos, data, 0.25)
    m = CudaSubVoxelMap.from_chunks(pos, data, 0.25)
    assert_same_map(ref, h, m)
    assert m.free_map().tobytes() == free_ref.tobytes()
    ref.map_destroy(h)


@needs_ref
@pytest.mark.parametrize("kind", ["full range", "truncation boundary", "dense band"])
@pytest.mark.parametrize("pos", [[(-2, 1, 0), (3, -1, 2)], [(100, -100, 7)], [(-32, 5, 0), (31, 5, 0)]], ids=["apart", "far", "wide"])
def test_random_words_match_createTSDFMap(pos, kind):
    """Arbitrary TSDFValue words — int16 values over the whole range or sitting on the +-600 mm truncation bounds, negative
    weights — through both ingests: same geometry, brick table, payload and free-space points."""
    ref = Ref()
    rng = np.random.default_rng(len(pos) + len(kind))
    shape = (len(pos), 64, 64, 64)
    if kind == "full range":
        value = rng.integers(-32768, 32768, size=shape)
        weight = np.where(rng.random(shape) < 0.02, rng.integers(-32768, 32768, size=shape), 0)
    elif kind == "truncation boundary":
        value = rng.choice([-601, -600, -599, -1, 0, 1, 599, 600, 601], size=shape)
        weight = np.where(rng.random(shape) < 0.05, rng.integers(1, 5, size=shape), 0)
    else:
        value = rng.integers(-700, 701, size=shape)
        weight = np.where(rng.random(shape) < 0.3, 1, 0)
    data = value.astype(np.int16).view(np.uint16).astype(np.uint32) | (weight.astype(np.int16).view(np.uint16).astype(np.uint32) << 16)
    h, free_ref = ref.create_tsdf_map(pos, data, 0.1)
    m = CudaSubVoxelMap.from_chunks(pos, data, 0.1)
    assert_same_map(ref, h, m)
    assert m.free_map().tobytes() == free_ref.tobytes()
    ref.map_destroy(h)


def test_rejects_bad_input():
    with pytest.raises(ValueError):
        CudaSubVoxelMap.from_chunks([(0, 0, 0)], np.zeros((1, 10), dtype=np.uint32))
    with pytest.raises(ValueError):
        CudaSubVoxelMap.from_chunks([(0, 0, 0), (0, 0, 0)], np.zeros((2, 64 ** 3), dtype=np.uint32))      # duplicate chunk
    with pytest.raises(ValueError):
        CudaSubVoxelMap.from_chunks([(0, 0, 0)], np.zeros((1, 64 ** 3), dtype=np.uint32), sigma=0.0)


def test_gpu_ingest_fails_loudly_without_gpu():
    """No silent fallback: asking for the GPU ingest on a machine without a CUDA device is an error, not a host computation."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        CudaSubVoxelMap.from_chunks([(0, 0, 0)], np.zeros((1, 64 ** 3), dtype=np.uint32), 0.1, device=0)
