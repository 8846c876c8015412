#!/bin/bash
# usage: scripts/sweep.sh N KIND "ENV1" "ENV2" ...   — one quick_gpu.py run per environment string
N=$1; KIND=$2; shift 2
for e in "$@"; do
  echo "== $e"
  env $e timeout 300 python scripts/quick_gpu.py $N $KIND 2 2>&1 | tail -3
done
