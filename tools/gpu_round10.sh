#!/bin/bash
set -u
TAG=${1:-run10}
mkdir -p gpurun_out
for v in f2_13 f2_12 f2_35 f2_23 f2_34 f2_11; do
  echo "variant $v" >> gpurun_out/${TAG}_dp_variants.jsonl
  OVL_B200_LIB=build/variants/libovl_$v.so python tools/dp_sweep.py --only 4x38 --modes 1 --reps 5 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
  OVL_B200_LIB=build/variants/libovl_$v.so python tools/dp_sweep.py --workload ecoli_n200k_l1000 --k 8 --only 32x32 --modes 1 --reps 3 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
done
grep -v '"lib"' gpurun_out/${TAG}_dp_variants.jsonl
tail -3 gpurun_out/${TAG}_dp.err
