#!/bin/bash
# k-mer stage profiles + ncu launch list of a short bench run
set -u
TAG=${1:-run3}
mkdir -p gpurun_out
for cfg in "1000000 5" "1000000 10" "1000000 15" "8000000 8" "8000000 10" "8000000 15"; do
  python tools/kmer_profile.py $cfg >> gpurun_out/${TAG}_kmer_profile.jsonl 2>> gpurun_out/${TAG}_kmer_profile.err
done
cat gpurun_out/${TAG}_kmer_profile.jsonl
tail -3 gpurun_out/${TAG}_kmer_profile.err
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench_short.json 2> gpurun_out/${TAG}_bench_short.err
echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_phix50k.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --workload phix_n50000_l150 > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu2 rc=$?"
