set -x
python -m pytest tests/test_gpu_graph.py -x -q -m gpu 2>&1 | tail -15
for w in c1 c2; do for g in 1 0; do
python bench.py --workload $w --graphs $g --steps 50 --warmup 10 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02w_bench_${w}_graphs$g.json 2> gpurun_out/r02w_bench_${w}_graphs$g.err || tail -5 gpurun_out/r02w_bench_${w}_graphs$g.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02w_bench_${w}_graphs$g.json").read().strip().splitlines()[-1])
print("$w graphs=$g", d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["graphs"], d["gpu_launches"], d["config"]["weights_sha256"][:12], d["config"]["resampled_sha256"][:12], d["config"]["e2e_resampled_sha256"][:12])
PY
done; done
