set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py > gpurun_out/r02z_bench_c3.json 2> gpurun_out/r02z_bench_c3.err || tail -5 gpurun_out/r02z_bench_c3.err
python bench.py --workload c2 --steps 50 --warmup 10 > gpurun_out/r02z_bench_c2.json 2> gpurun_out/r02z_bench_c2.err || tail -5 gpurun_out/r02z_bench_c2.err
python bench.py --workload c1 --steps 50 --warmup 10 > gpurun_out/r02z_bench_c1.json 2> gpurun_out/r02z_bench_c1.err || tail -5 gpurun_out/r02z_bench_c1.err
python bench.py --workload c3 --particles 8192 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02z_bench_c3_8192.json 2> /dev/null
python - <<PY
import json
for w in ("c3","c2","c1","c3_8192"):
    d=json.loads(open(f"gpurun_out/r02z_bench_{w}.json").read().strip().splitlines()[-1])
    print(w, round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), "kernel", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],3), d["graphs"], d["gpu_launches"], d["config"]["weights_sha256"][:12], d["config"]["resampled_sha256"][:12], d["config"]["e2e_resampled_sha256"][:12], d.get("reference_cuda_ms_per_update"), (d.get("cpu_baseline") or {}).get("value"))
PY
