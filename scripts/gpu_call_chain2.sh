set -x
run() {  # name, args...
  name=$1; shift
  python bench.py "$@" --steps 20 --warmup 5 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02y_$name.json 2> gpurun_out/r02y_$name.err || tail -5 gpurun_out/r02y_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02y_$name.json").read().strip().splitlines()[-1])
print("$name", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), "kernel", round(d["roofline"]["kernel_ms"],4), d["config"]["weights_sha256"][:12], d["config"]["resampled_sha256"][:12])
PY
}
run c2_auto --workload c2
run c2_pair_auto --workload c2 --pairing 1
run c2_pair_c4 --workload c2 --pairing 1 --chunks 4
run c3_16384_shallow --workload c3 --particles 16384 --registers 1
run c3_16384_auto --workload c3 --particles 16384
run c3_12000_whole --workload c3 --particles 12000 --chunks 1
run c3_12000_auto --workload c3 --particles 12000
run c3_5000_whole --workload c3 --particles 5000 --chunks 1
run c3_5000_auto --workload c3 --particles 5000
run c3_5000_pair --workload c3 --particles 5000 --pairing 1
