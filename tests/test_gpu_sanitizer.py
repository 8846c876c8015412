"""compute-sanitizer memcheck over __graft_entry__.smoke() — the evaluation kernel, the cooperative normalisation / CDF
kernel, the draw kernel and the scan-reduction kernels, with the oracle checking every result — as a -m gpu test
(VERDICT r01 Next #10). Skipped when the tool is not installed on the box."""
import os
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]


def _tool():
    for cand in (shutil.which("compute-sanitizer"), "/usr/local/cuda/bin/compute-sanitizer"):
        if cand and Path(cand).exists():
            return cand
    return None


@pytest.mark.parametrize("tool", ["memcheck", "racecheck"])
def test_smoke_is_clean_under_compute_sanitizer(tool):
    exe = _tool()
    if exe is None:
        pytest.skip("compute-sanitizer is not installed")
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    cmd = [exe, "--tool", tool, "--error-exitcode", "86", sys.executable, "-c",
           "import __graft_entry__ as g; g.smoke()"]
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (res.stdout + res.stderr)[-3000:]
    assert res.returncode == 0, tail
    assert "smoke ok" in res.stdout, tail
    if tool == "memcheck":
        assert "ERROR SUMMARY: 0 errors" in res.stdout + res.stderr, tail
    else:
        assert "RACECHECK SUMMARY: 0 hazards" in res.stdout + res.stderr, tail
