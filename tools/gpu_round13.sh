#!/bin/bash
# DP column-form variants (two-input packed min forms 3/4/5): kernel timed alone on the PhiX-50k and 200k x 1000 bp pair lists
set -u
TAG=${1:-run13}
shift
mkdir -p gpurun_out
for v in "$@"; do
  echo "variant $v" >> gpurun_out/${TAG}_dp_variants.jsonl
  OVL_B200_LIB=build/variants/libovl_$v.so python tools/dp_sweep.py --only 4x38 --modes 1 --reps 5 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
  OVL_B200_LIB=build/variants/libovl_$v.so python tools/dp_sweep.py --workload ecoli_n200k_l1000 --k 8 --only 32x32 --modes 1 --reps 3 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
done
grep -v '"lib"' gpurun_out/${TAG}_dp_variants.jsonl
tail -3 gpurun_out/${TAG}_dp.err
# probes of the newest kinds, from the last variant's library
OVL_B200_LIB=build/variants/libovl_${v}.so python tools/dp_sweep.py --probe-only --max-pairs 1000 >> gpurun_out/${TAG}_probe.jsonl 2>> gpurun_out/${TAG}_dp.err
tail -1 gpurun_out/${TAG}_probe.jsonl
