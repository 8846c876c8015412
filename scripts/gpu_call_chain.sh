set -x
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_graph.py -x -q -m gpu 2>&1 | tail -15
run() {  # name, args...
  name=$1; shift
  python bench.py "$@" --steps 20 --warmup 5 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02x_$name.json 2> gpurun_out/r02x_$name.err || tail -5 gpurun_out/r02x_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02x_$name.json").read().strip().splitlines()[-1])
print("$name", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), "kernel", round(d["roofline"]["kernel_ms"],4), d["config"]["weights_sha256"][:12], d["config"]["resampled_sha256"][:12])
PY
}
for c in 1 0 4 16; do run c3_8192_chunks$c --workload c3 --particles 8192 --chunks $c; done
for c in 1 0 4; do run c2_chunks$c --workload c2 --chunks $c; done
for c in 1 0; do run c3_16384_chunks$c --workload c3 --particles 16384 --chunks $c; done
for c in 1 0; do run c3_32768_chunks$c --workload c3 --particles 32768 --chunks $c; done
