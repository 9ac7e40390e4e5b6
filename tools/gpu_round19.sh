#!/bin/bash
# L2 policy experiments for the DP kernel: time (kernel alone) and DRAM traffic (one bench step under ncu) per variant library
set -u
TAG=${1:-run19}
shift
mkdir -p gpurun_out
for v in "$@"; do
  echo "variant $v" >> gpurun_out/${TAG}_dp_variants.jsonl
  OVL_B200_LIB=build/variants/libovl_$v.so python tools/dp_sweep.py --only 4x38 --reps 7 --modes 1 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
  OVL_B200_LIB=build/variants/libovl_$v.so ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:overlap_dp_kernel -c 1 --csv \
    --log-file gpurun_out/${TAG}_dp_traffic_$v.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_ncu_traffic_$v.log 2>&1
  echo "$v ncu rc=$?"; tail -3 gpurun_out/${TAG}_dp_traffic_$v.csv | awk -F'","' '{print $13, $15}'
done
grep -v '"lib"' gpurun_out/${TAG}_dp_variants.jsonl
tail -3 gpurun_out/${TAG}_dp.err
