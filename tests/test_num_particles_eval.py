"""The reference's OWN benchmark program, src/num_particles_eval.cpp, compiled unmodified (oracle/Makefile `npe`,
oracle/npe_harness.cpp: ROS parameters from the environment, the HDF5 map file from a raw chunk dump) and run end to end:
snapshot -> scan reduction -> createTSDFMap -> TSDFEvaluator -> timed evaluate() sweep. CPU: the use_cuda=false build. -m gpu:
the build linked against the product's drop-in CudaEvaluator — the reference's benchmark running on the B200 unchanged."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "oracle" / "_ref"


def run(impl, n=300, inc=100, repeat=1):
    res = subprocess.run([sys.executable, str(ROOT / "scripts" / "run_num_particles_eval.py"), "--impl", impl, "--num-particles", str(n),
                          "--inc", str(inc), "--repeat", str(repeat), "--timeout", "120"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    return [json.loads(line) for line in res.stdout.splitlines() if line.startswith("{")]


@pytest.mark.skipif(not (EXE / "num_particles_eval_cpu").exists(), reason="oracle/_ref/num_particles_eval_cpu not built")
def test_reference_benchmark_program_runs_on_its_cpu_evaluator():
    (r,) = run("cpu")
    assert r["rc"] == 0 and r["finished"], r
    assert [row[0] for row in r["num_particles__runtime_ms"]] == [100, 200, 300]
    assert 0 < int(r["scan_points_evaluated"]) < r["snapshot_points"]          # the program's own 6.4 cm reduction ran


@pytest.mark.skipif(not (EXE / "num_particles_eval_b200").exists(), reason="oracle/_ref/num_particles_eval_b200 not built")
def test_drop_in_build_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    (r,) = run("b200")
    assert r["rc"] == 1 and not r["finished"]
    assert "no CPU fallback" in " ".join(r["stderr_tail"])


@pytest.mark.gpu
@pytest.mark.skipif(not (EXE / "num_particles_eval_b200").exists(), reason="oracle/_ref/num_particles_eval_b200 not built")
def test_reference_benchmark_program_runs_on_the_b200_through_the_drop_in():
    (r,) = run("b200", n=60000, inc=20000, repeat=2)
    assert r["rc"] == 0 and r["finished"], r
    assert [row[0] for row in r["num_particles__runtime_ms"]] == [20000, 40000, 60000]
