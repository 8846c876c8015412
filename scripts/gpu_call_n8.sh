timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02b_bench_c3_n8.json 2> gpurun_out/r02b_bench_c3_n8.err || tail -5 gpurun_out/r02b_bench_c3_n8.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02b_bench_c3_n8.json").read().strip().splitlines()[-1])
print("n8", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), d["e2e"].get("per_process_ms_per_step"), "kernel", round(d["roofline"]["kernel_ms"],4), d["config"]["weights_sha256"][:12], d["config"]["resampled_sha256"][:12], d["config"]["e2e_resampled_sha256"][:12], d["config"]["transport"])
PY
