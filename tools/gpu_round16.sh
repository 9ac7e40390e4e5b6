#!/bin/bash
# the full GPU suite on the default library, then the DP kernel alone (4x38 on the PhiX-50k pairs) for every named variant + default
set -u
TAG=${1:-run16}
shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
for v in "$@" default; do
  echo "variant $v" >> gpurun_out/${TAG}_dp_variants.jsonl
  if [ "$v" = default ]; then unset OVL_B200_LIB; else export OVL_B200_LIB=build/variants/libovl_$v.so; fi
  python tools/dp_sweep.py --only 4x38 --modes 1 --reps 7 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
done
unset OVL_B200_LIB
python tools/dp_sweep.py --workload ecoli_n200k_l1000 --k 8 --only 32x32 --modes 1 --reps 3 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
grep -v '"lib"' gpurun_out/${TAG}_dp_variants.jsonl
tail -3 gpurun_out/${TAG}_dp.err
