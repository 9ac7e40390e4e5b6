#!/bin/bash
set -u
TAG=${1:-run8}
mkdir -p gpurun_out
for v in minb4 minb4_f2_14 minb4_f2_12 minb4_f2_25 minb5 minb5_f2_12; do
  echo "variant $v" >> gpurun_out/${TAG}_dp_variants.jsonl
  OVL_B200_LIB=build/variants/libovl_$v.so python tools/dp_sweep.py --only 4x38 --modes 1 --reps 5 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
  OVL_B200_LIB=build/variants/libovl_$v.so python tools/dp_sweep.py --workload ecoli_n200k_l1000 --k 8 --only 32x32 --modes 1 --reps 3 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
  OVL_B200_LIB=build/variants/libovl_$v.so python tools/dp_sweep.py --workload phix_n1000_l100 --only 4x25 --modes 1 --reps 5 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
done
grep -v '"lib"' gpurun_out/${TAG}_dp_variants.jsonl
tail -3 gpurun_out/${TAG}_dp.err
