"""tsdf_localization_b200 — B200-native (sm_100a) MCL sensor update of uos/tsdf_localization.

Only the hot path lives here: the CUDA kernels + C ABI (``csrc/`` -> ``lib/libtsdfloc.so``) and the host-side
mirror of the reference's evaluator / resampler interface (``evaluator.py``), plus the synthetic workload
generators the tests and ``bench.py`` share (``synthetic.py``) and the one-process-per-GPU driver (``dist.py``).
There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
from .capi import LibraryNotBuilt, TsdflocError, lib_path, load_library  # noqa: F401
from .evaluator import (CudaEvaluator, CudaSubVoxelMap, DrawSource, MetropolisResampler, MultiGpuEvaluator, ParticleCloud,  # noqa: F401
                        RejectionResampler, ResidualResampler, ResidualSystematicResampler, SystematicResampler, TSDFEvaluator,
                        WheelResampler, likelihood_init, likelihood_value)

__all__ = [
    "CudaEvaluator", "CudaSubVoxelMap", "DrawSource", "MetropolisResampler", "MultiGpuEvaluator", "ParticleCloud", "RejectionResampler",
    "ResidualResampler", "ResidualSystematicResampler", "SystematicResampler", "TSDFEvaluator", "WheelResampler", "likelihood_init",
    "likelihood_value",
    "LibraryNotBuilt", "TsdflocError", "lib_path", "load_library",
]
