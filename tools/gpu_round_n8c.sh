#!/bin/bash
# 8-GPU box: 2-GPU hardware parity test, then the N = 8 (and optionally N = 4, 2) bench lines
set -u
TAG=${1:-n8c}
shift
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/${TAG}_pytest_multi.log 2>&1
echo "pytest multi rc=$?"; tail -5 gpurun_out/${TAG}_pytest_multi.log
for N in "$@"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
      bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/${TAG}_bench_n$N.err | grep '^{"metric"' > gpurun_out/${TAG}_bench_n$N.json
  echo "bench N=$N rc=$?"; tail -c 300 gpurun_out/${TAG}_bench_n$N.err
  python - $TAG $N <<'PY'
import json,sys
d=json.load(open("gpurun_out/%s_bench_n%s.json" % (sys.argv[1], sys.argv[2])))
print({k:d[k] for k in ("value","ms_per_step","e2e")})
PY
done
