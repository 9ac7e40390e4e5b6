#!/bin/bash
set -u
TAG=${1:-run12}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
for cfg in "1000000 5" "1000000 10" "8000000 8" "8000000 10" "8000000 15"; do
  python tools/kmer_profile.py $cfg >> gpurun_out/${TAG}_kmer_profile.jsonl 2>> gpurun_out/${TAG}_kmer_profile.err
done
tail -3 gpurun_out/${TAG}_kmer_profile.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_kmer_kernels_8m_k8.csv python tools/kmer_profile.py 8000000 8 > gpurun_out/${TAG}_ncu_kmer.log 2>&1
echo "ncu kmer rc=$?"
