#!/bin/bash
# full GPU suite on the default library; DP kernel alone: default (4x38, 32x32), phases (4x38), rowall (32x32); DP DRAM traffic of one bench step
set -u
TAG=${1:-run17}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
run() { # variant workload-args...
  v=$1; shift
  echo "variant $v" >> gpurun_out/${TAG}_dp_variants.jsonl
  if [ "$v" = default ]; then unset OVL_B200_LIB; else export OVL_B200_LIB=build/variants/libovl_$v.so; fi
  python tools/dp_sweep.py "$@" --modes 1 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
  unset OVL_B200_LIB
}
run default --only 4x38 --reps 7
run phases --only 4x38 --reps 7
run default --workload ecoli_n200k_l1000 --k 8 --only 32x32 --reps 3
run rowall --workload ecoli_n200k_l1000 --k 8 --only 32x32 --reps 3
run nobulk --workload ecoli_n200k_l1000 --k 8 --only 32x32 --reps 3
grep -v '"lib"' gpurun_out/${TAG}_dp_variants.jsonl
tail -3 gpurun_out/${TAG}_dp.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:overlap_dp_kernel -c 1 --csv \
    --log-file gpurun_out/${TAG}_dp_traffic.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"; tail -4 gpurun_out/${TAG}_dp_traffic.csv | cut -c1-400
