#!/bin/bash
# GPU tests, then the DP kernel A/B (register vs immediate gap costs) + integer probes
set -u
TAG=${1:-run6}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
OVL_DP_IMM_GAPS=0 python tools/dp_sweep.py --only 4x38 --modes 1 --reps 5 --no-probe > gpurun_out/${TAG}_dp_reg.jsonl 2> gpurun_out/${TAG}_dp.err
OVL_DP_IMM_GAPS=1 python tools/dp_sweep.py --only 4x38 --modes 1 --reps 5 > gpurun_out/${TAG}_dp_imm.jsonl 2>> gpurun_out/${TAG}_dp.err
OVL_DP_IMM_GAPS=0 python tools/dp_sweep.py --workload ecoli_n200k_l1000 --k 8 --only 32x32 --modes 1 --reps 3 --no-probe > gpurun_out/${TAG}_dp_reg_l1000.jsonl 2>> gpurun_out/${TAG}_dp.err
OVL_DP_IMM_GAPS=1 python tools/dp_sweep.py --workload ecoli_n200k_l1000 --k 8 --only 32x32 --modes 1 --reps 3 --no-probe > gpurun_out/${TAG}_dp_imm_l1000.jsonl 2>> gpurun_out/${TAG}_dp.err
cat gpurun_out/${TAG}_dp_reg.jsonl gpurun_out/${TAG}_dp_imm.jsonl gpurun_out/${TAG}_dp_reg_l1000.jsonl gpurun_out/${TAG}_dp_imm_l1000.jsonl
tail -3 gpurun_out/${TAG}_dp.err
