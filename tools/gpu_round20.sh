#!/bin/bash
# full GPU suite + k-mer stage profile (default library)
set -u
TAG=${1:-run20}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
for cfg in "1000000 5" "1000000 10" "8000000 8" "8000000 10" "8000000 15"; do
  python tools/kmer_profile.py $cfg >> gpurun_out/${TAG}_kmer_profile.jsonl 2>> gpurun_out/${TAG}_kmer_profile.err
done
tail -3 gpurun_out/${TAG}_kmer_profile.err
python - $TAG <<'PY'
import json,sys
for ln in open("gpurun_out/%s_kmer_profile.jsonl" % sys.argv[1]):
    d=json.loads(ln)
    print(d["reads"], d["k"], {k[:13]: (round(v["ms"]*1000), round(v["frac_of_hbm_peak"],2)) for k,v in d["stages"].items()}, "sum", round(d["sum"]["ms"]*1000), round(d["sum"]["frac_of_hbm_peak"],3))
PY
