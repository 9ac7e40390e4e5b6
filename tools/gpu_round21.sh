#!/bin/bash
# sort tile size A/B: index/join parity tests + k-mer stage profile per variant library, then the full suite on the default library
set -u
TAG=${1:-run21}
shift
mkdir -p gpurun_out
for v in "$@"; do
  OVL_B200_LIB=build/variants/libovl_$v.so python -m pytest tests -m gpu -x -q -k "index or join or one_call or golden or config" > gpurun_out/${TAG}_pytest_$v.log 2>&1
  echo "pytest ($v) rc=$?"; tail -2 gpurun_out/${TAG}_pytest_$v.log
  echo "variant $v" >> gpurun_out/${TAG}_kmer_profile.jsonl
  for cfg in "8000000 8" "8000000 15" "1000000 10"; do
    OVL_B200_LIB=build/variants/libovl_$v.so python tools/kmer_profile.py $cfg >> gpurun_out/${TAG}_kmer_profile.jsonl 2>> gpurun_out/${TAG}_kmer_profile.err
  done
done
true
true
tail -3 gpurun_out/${TAG}_kmer_profile.err
python - $TAG <<'PY'
import json,sys
for ln in open("gpurun_out/%s_kmer_profile.jsonl" % sys.argv[1]):
    if ln.startswith("variant"): print(ln.strip()); continue
    d=json.loads(ln)
    print(d["reads"], d["k"], {k[:13]: (round(v["ms"]*1000), round(v["frac_of_hbm_peak"],2)) for k,v in d["stages"].items()}, "sum", round(d["sum"]["ms"]*1000), round(d["sum"]["frac_of_hbm_peak"],3))
PY
