#!/bin/bash
set -u
TAG=${1:-run11}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "one_call or keys_index or golden or duplicates" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
ncu --set full --clock-control none --import-source on -k regex:sort_scatter_kernel -s 4 -c 2 -o gpurun_out/${TAG}_scatter_full \
    python tools/kmer_profile.py 8000000 8 > gpurun_out/${TAG}_ncu_scatter.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${TAG}_ncu_scatter.log
python tools/kmer_profile.py 1000000 5 > gpurun_out/${TAG}_kmer_profile.jsonl 2>&1
