#!/bin/bash
# GPU tests + k-mer stage profile + a short bench line (no extras)
set -u
TAG=${1:-run4}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
for cfg in "1000000 5" "1000000 10" "8000000 8" "8000000 10" "8000000 15"; do
  python tools/kmer_profile.py $cfg >> gpurun_out/${TAG}_kmer_profile.jsonl 2>> gpurun_out/${TAG}_kmer_profile.err
done
tail -3 gpurun_out/${TAG}_kmer_profile.err
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench_short.json 2> gpurun_out/${TAG}_bench_short.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_short.err
python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --workload phix_n50000_l150 > gpurun_out/${TAG}_bench_phix50k.json 2>> gpurun_out/${TAG}_bench_short.err
python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --workload phix_n1000_l100 > gpurun_out/${TAG}_bench_phix1k.json 2>> gpurun_out/${TAG}_bench_short.err
