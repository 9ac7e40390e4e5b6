#!/bin/bash
# pack kernel preload-count A/B at 8 M reads (+ default)
set -u
TAG=${1:-run24}
shift
mkdir -p gpurun_out
for v in "$@" default; do
  echo "variant $v" >> gpurun_out/${TAG}_kmer_profile.jsonl
  if [ "$v" = default ]; then unset OVL_B200_LIB; else export OVL_B200_LIB=build/variants/libovl_$v.so; fi
  python tools/kmer_profile.py 8000000 8 >> gpurun_out/${TAG}_kmer_profile.jsonl 2>> gpurun_out/${TAG}_kmer_profile.err
done
unset OVL_B200_LIB
tail -2 gpurun_out/${TAG}_kmer_profile.err
python - $TAG <<'PY'
import json,sys
for ln in open("gpurun_out/%s_kmer_profile.jsonl" % sys.argv[1]):
    if ln.startswith("variant"): print(ln.strip()); continue
    d=json.loads(ln)
    print(d["reads"], d["k"], {k[:13]: (round(v["ms"]*1000,1), round(v["frac_of_hbm_peak"],3)) for k,v in d["stages"].items()})
PY
