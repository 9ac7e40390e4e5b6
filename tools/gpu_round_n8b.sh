#!/bin/bash
# 8-GPU box: aggregate device->host ceiling (own buffers / shared segment / NUMA-local shared segment), then the N = 8 bench line
set -u
TAG=${1:-n8b}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29617 \
    tools/d2h_ceiling.py 2> gpurun_out/${TAG}_d2h.err | grep '^{' > gpurun_out/${TAG}_d2h_ceiling.json
echo "d2h rc=$?"; cat gpurun_out/${TAG}_d2h_ceiling.json; tail -c 300 gpurun_out/${TAG}_d2h.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 8 --steps 5 --warmup 3 2> gpurun_out/${TAG}_bench_n8.err | grep '^{"metric"' > gpurun_out/${TAG}_bench_n8.json
echo "bench N=8 rc=$?"; tail -c 400 gpurun_out/${TAG}_bench_n8.err
python - $TAG <<'PY'
import json,sys
d=json.load(open("gpurun_out/%s_bench_n8.json" % sys.argv[1] if len(sys.argv)>1 else "gpurun_out/n8b_bench_n8.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e")})
PY
