#!/bin/bash
# GPU tests, k-mer stage profile (+ per-kernel ncu list at 8 M reads), integer probes, one full ncu capture of the DP kernel
set -u
TAG=${1:-run5}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
for cfg in "1000000 5" "1000000 10" "8000000 8" "8000000 10" "8000000 15"; do
  python tools/kmer_profile.py $cfg >> gpurun_out/${TAG}_kmer_profile.jsonl 2>> gpurun_out/${TAG}_kmer_profile.err
done
tail -3 gpurun_out/${TAG}_kmer_profile.err
python tools/dp_sweep.py --only 4x38 --modes 1 --reps 3 > gpurun_out/${TAG}_dp_sweep.jsonl 2> gpurun_out/${TAG}_dp_sweep.err
echo "sweep rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_kmer_kernels_8m_k8.csv python tools/kmer_profile.py 8000000 8 > gpurun_out/${TAG}_ncu_kmer.log 2>&1
echo "ncu kmer rc=$?"
ncu --set full --clock-control none --import-source on -k regex:overlap_dp_kernel -c 1 -o gpurun_out/${TAG}_dp_full \
    python tools/dp_sweep.py --only 4x38 --modes 1 --no-probe --reps 1 > gpurun_out/${TAG}_ncu_dp.log 2>&1
echo "ncu dp rc=$?"
ls -la gpurun_out | tail -12
