#!/bin/bash
# final validation: GPU suite, default bench (with configs_extra), launch list of one bench step
set -u
TAG=${1:-run22}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/${TAG}_launches_ecoli1m.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu bench rc=$?"
python - $TAG <<'PY'
import json,sys
d=json.load(open("gpurun_out/%s_bench.json" % sys.argv[1]))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["kmer_stages"]["ms"], d["kmer_stages"]["k0_k3_frac"])
n=d["configs_extra"]["next_rows"]; print(n.get("error"), n.get("f1_local_alignment",{}).get("batch_ms_per_contig"))
for k in ("configs[0]","configs[1]","configs[3]_k8","configs[3]_k5"):
    v=d["configs_extra"][k]; print(k, v.get("error"), round(v["value"],1), round(v["e2e"]["value"],1), round(v["stage_ms"]["kmer_index_join"],3))
PY
