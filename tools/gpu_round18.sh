#!/bin/bash
# DP kernel alone, A/B of code shapes: head (committed), lambda step with/without streaming hints, phases, rowall
set -u
TAG=${1:-run18}
mkdir -p gpurun_out
run() { v=$1; shift
  echo "variant $v" >> gpurun_out/${TAG}_dp_variants.jsonl
  OVL_B200_LIB=build/variants/libovl_$v.so python tools/dp_sweep.py "$@" --modes 1 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
}
for v in head lam_nohints lam_hints phases_nohints; do run $v --only 4x38 --reps 7; done
for v in head lam_nohints rowall_nohints lam_hints; do run $v --workload ecoli_n200k_l1000 --k 8 --only 32x32 --reps 3; done
grep -v '"lib"' gpurun_out/${TAG}_dp_variants.jsonl
tail -3 gpurun_out/${TAG}_dp.err
