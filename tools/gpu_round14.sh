#!/bin/bash
# full GPU suite; pack-kernel occupancy A/B (base = previous pack + per-read finalize); default bench with next_rows
set -u
TAG=${1:-run14}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
for v in base pack4 pack5 pack8; do
  echo "variant $v" >> gpurun_out/${TAG}_kmer_profile.jsonl
  OVL_B200_LIB=build/variants/libovl_$v.so python tools/kmer_profile.py 8000000 8 >> gpurun_out/${TAG}_kmer_profile.jsonl 2>> gpurun_out/${TAG}_kmer_profile.err
done
echo "variant default (pack6)" >> gpurun_out/${TAG}_kmer_profile.jsonl
for cfg in "8000000 8" "1000000 5" "1000000 10"; do
  python tools/kmer_profile.py $cfg >> gpurun_out/${TAG}_kmer_profile.jsonl 2>> gpurun_out/${TAG}_kmer_profile.err
done
tail -3 gpurun_out/${TAG}_kmer_profile.err
python - $TAG <<'PY'
import json,sys
for ln in open("gpurun_out/%s_kmer_profile.jsonl" % sys.argv[1]):
    if ln.startswith("variant"): print(ln.strip()); continue
    d=json.loads(ln)
    print(d["reads"], d["k"], {k[:8]: round(v["ms"]*1000) for k,v in d["stages"].items()}, "sum", round(d["sum"]["ms"]*1000), round(d["sum"]["frac_of_hbm_peak"],3))
PY
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${TAG}_bench.err
python - $TAG <<'PY'
import json,sys
d=json.load(open("gpurun_out/%s_bench.json" % sys.argv[1]))
print({k:d[k] for k in ("value","ms_per_step","e2e","kmer_stages","roofline")})
print(json.dumps(d["configs_extra"].get("next_rows"))[:3000])
PY
