#!/bin/bash
# A/B on ONE box, alternating (box-to-box clock / power-state differences are +-3 %, larger than most effects):
# usage  ab_bench.sh VAR "0 1" [rounds] [pairs per step] [steps]
VAR=$1; VALS=$2; ROUNDS=${3:-2}; PAIRS=${4:-1}; STEPS=${5:-10}
mkdir -p gpurun_out
for r in $(seq 1 $ROUNDS); do
  for v in $VALS; do
    env $VAR=$v timeout 300 python bench.py --steps $STEPS --warmup 3 --pairs $PAIRS --quick --no-cpu-baseline > gpurun_out/ab_${VAR}_${v}_$r.json 2>> gpurun_out/ab.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/ab_${VAR}_${v}_$r.json").read().strip().splitlines()[-1])
k = d["kernel_breakdown_ms"]
print("$VAR=$v pairs=$PAIRS round $r: %.2f ms/step  %.3f pairs/s  clock %s  fmha %.1f  gemm %.1f  ln %.1f" % (d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"], k.get("fmha_varlen", 0),
      sum(x for n, x in k.items() if n.startswith("gemm")), k.get("ln_modulate", 0)))
PY
  done
done
