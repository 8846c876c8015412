set -x
NCU="ncu --set full --clock-control none --import-source on -k regex:k_eval -s 1 -c 1 -f"
timeout 400 $NCU -o gpurun_out/r02b_ncu_8192_chain python scripts/profile_eval.py 8192 os1-128 2 2>&1 | tail -3
timeout 400 $NCU -o gpurun_out/r02b_ncu_c2_chain python scripts/profile_eval.py 5000 vlp16 2 2>&1 | tail -3
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r02b_launches_bench_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > /dev/null 2>&1
for f in r02b_ncu_8192_chain r02b_ncu_c2_chain; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv; done
ls -la gpurun_out | tail -8
