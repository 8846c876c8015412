"""In-tree builds: libtsdfloc.so (nvcc, sm_100a) and the CPU checkers under oracle/ (make).

Used by ``__graft_entry__.build()``; also runnable as ``python -m tsdf_localization_b200.build``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libtsdfloc.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build libtsdfloc.so")


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build_variant(name: str, defs, verbose: bool = False) -> Path:
    """Tuning experiments: lib/libtsdfloc_<name>.so built with extra -D flags (select with TSDFLOC_LIB=<path>)."""
    out = LIB.parent / f"libtsdfloc_{name}.so"
    sources = [CSRC / "tsdfloc_api.cu", CSRC / "host_map.cpp", CSRC / "host_motion.cpp", CSRC / "host_resample.cpp"]
    cmd = [_nvcc(), "-ccbin", "/usr/bin/g++", *NVCC_FLAGS, *defs, *(["-Xptxas", "-v"] if verbose else []), "-o", str(out), *map(str, sources)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source of the package for sm_100a into lib/libtsdfloc.so."""
    sources = [CSRC / "tsdfloc_api.cu", CSRC / "host_map.cpp", CSRC / "host_motion.cpp", CSRC / "host_resample.cpp"]
    deps = sources + [CSRC / "tsdfloc_kernels.cuh", CSRC / "tsdfloc_eval.cuh", CSRC / "tsdfloc_device.cuh", CSRC / "tsdfloc_reduce.cuh",
                      CSRC / "tsdfloc_motion.cuh", CSRC / "tsdfloc_sort.cuh", CSRC / "tsdfloc_multi.inc", CSRC / "tsdfloc_ingest.inc", CSRC / "tsdfloc_graph.inc", CSRC / "tsdfloc_mcl.inc", CSRC / "tsdfloc_host_map.h", ROOT / "include" / "tsdfloc.h"]
    if not force and not _stale(LIB, deps):
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    extra = os.environ.get("TSDFLOC_NVCC_DEFS", "").split()   # e.g. "-DTSDFLOC_BLOCK_STEPS=8" for tuning experiments
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, *(["-Xptxas", "-v"] if verbose else []), "-o", str(LIB), *map(str, sources)]
    env = dict(os.environ)
    # the image's CC/CXX wrappers point at a gcc without libgomp specs; nvcc is fine with the system one
    if Path("/usr/bin/g++").exists():
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


def build_oracle(verbose: bool = False) -> None:
    """Build oracle/_build/libtsdf_oracle.so and, when /root/reference is present, oracle/_ref/*.so."""
    res = subprocess.run(["make", "-j8", "-C", str(ROOT / "oracle"), "all"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose=True)
    build_oracle(verbose=True)
    print("built", LIB)
