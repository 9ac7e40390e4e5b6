#!/bin/bash
# DP kernel A/B: the full GPU suite under the FIRST named variant library, then the kernel alone for every variant and the default build
set -u
TAG=${1:-run15}
shift
mkdir -p gpurun_out
OVL_B200_LIB=build/variants/libovl_$1.so python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_$1.log 2>&1
echo "pytest ($1) rc=$?"; tail -4 gpurun_out/${TAG}_pytest_$1.log
for v in "$@" default; do
  echo "variant $v" >> gpurun_out/${TAG}_dp_variants.jsonl
  if [ "$v" = default ]; then unset OVL_B200_LIB; else export OVL_B200_LIB=build/variants/libovl_$v.so; fi
  python tools/dp_sweep.py --only 4x38 --modes 1 --reps 5 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
  python tools/dp_sweep.py --workload ecoli_n200k_l1000 --k 8 --only 32x32 --modes 1 --reps 3 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
  python tools/dp_sweep.py --workload phix_n1000_l100 --only 4x25 --modes 1 --reps 5 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
done
grep -v '"lib"' gpurun_out/${TAG}_dp_variants.jsonl
tail -3 gpurun_out/${TAG}_dp.err
