#!/bin/bash
# Dev probe: tuning builds of libtsdfloc (lib/libtsdfloc_<name>.so from tsdf_localization_b200.build.build_variant) through
# the evaluation-kernel sweep.   usage: COUNTS=8192,65536 scripts/probes/variants.sh "" b4 b6
for lib in "$@"; do
  echo "== ${lib:-shipped}"
  if [ -n "$lib" ]; then export TSDFLOC_LIB=$PWD/tsdf_localization_b200/lib/libtsdfloc_$lib.so; else unset TSDFLOC_LIB; fi
  python scripts/sweep_eval.py gpurun_out/variant_${lib:-shipped}.jsonl os1-128 ${COUNTS:-500,8192,65536} 2>&1 | grep "'particles'" | python -c "
import sys,re
for l in sys.stdin:
    m=re.search(r\"'particles': (\d+), 'points': (\d+), 'pairing': '(\w+)', 'registers': (\d+), 'ms_min': ([\d.]+)\", l)
    if m: print(f'  {m.group(1):>6} x {m.group(2):>6} {m.group(3):9s} {m.group(4):>3} regs {float(m.group(5)):.3f} ms')
"
done
