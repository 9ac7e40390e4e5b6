#!/bin/bash
# DP variant sweep (immediate gap costs on) + k-mer profile + fill lanes A/B
set -u
TAG=${1:-run7}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "one_call or pack or keys_index or golden or batch_dp_lengths" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
for v in f2_0 f2_14 f2_12 f2_23 f2_11 minb4; do
  echo "variant $v" >> gpurun_out/${TAG}_dp_variants.jsonl
  OVL_B200_LIB=build/variants/libovl_$v.so python tools/dp_sweep.py --only 4x38 --modes 1 --reps 5 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
  OVL_B200_LIB=build/variants/libovl_$v.so python tools/dp_sweep.py --workload ecoli_n200k_l1000 --k 8 --only 32x32 --modes 1 --reps 3 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
done
echo "variant base" >> gpurun_out/${TAG}_dp_variants.jsonl
python tools/dp_sweep.py --only 4x38 --modes 1 --reps 5 --no-probe >> gpurun_out/${TAG}_dp_variants.jsonl 2>> gpurun_out/${TAG}_dp.err
grep -v '"lib"' gpurun_out/${TAG}_dp_variants.jsonl
tail -3 gpurun_out/${TAG}_dp.err
for cfg in "1000000 5" "1000000 10" "8000000 8" "8000000 10"; do
  python tools/kmer_profile.py $cfg >> gpurun_out/${TAG}_kmer_profile.jsonl 2>> gpurun_out/${TAG}_kmer_profile.err
done
for L in 8 4; do
  echo "fill lanes $L" >> gpurun_out/${TAG}_kmer_profile.jsonl
  OVL_FILL_LANES=$L python tools/kmer_profile.py 8000000 8 >> gpurun_out/${TAG}_kmer_profile.jsonl 2>> gpurun_out/${TAG}_kmer_profile.err
done
echo "fill lanes 8 (k=10)" >> gpurun_out/${TAG}_kmer_profile.jsonl
OVL_FILL_LANES=8 python tools/kmer_profile.py 8000000 10 >> gpurun_out/${TAG}_kmer_profile.jsonl 2>> gpurun_out/${TAG}_kmer_profile.err
echo "fill lanes 1 (k=10)" >> gpurun_out/${TAG}_kmer_profile.jsonl
OVL_FILL_LANES=1 python tools/kmer_profile.py 8000000 10 >> gpurun_out/${TAG}_kmer_profile.jsonl 2>> gpurun_out/${TAG}_kmer_profile.err
tail -3 gpurun_out/${TAG}_kmer_profile.err
