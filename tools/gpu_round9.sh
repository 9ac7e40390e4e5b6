#!/bin/bash
# full validation: GPU suite, default bench (with configs_extra), full ncu capture + launch list of the final kernels
set -u
TAG=${1:-run9}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench.err
ncu --set full --clock-control none --import-source on -k regex:overlap_dp_kernel -c 1 -o gpurun_out/${TAG}_dp_full \
    python tools/dp_sweep.py --only 4x38 --modes 1 --no-probe --reps 1 > gpurun_out/${TAG}_ncu_dp.log 2>&1
echo "ncu dp rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/${TAG}_launches_ecoli1m.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu bench rc=$?"
head -c 2500 gpurun_out/${TAG}_bench.json
