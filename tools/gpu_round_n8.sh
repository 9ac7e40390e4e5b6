#!/bin/bash
# 8-GPU box: bench lines at N = 8 and N = 4 (max-over-ranks device time), plus the 2-GPU parity test
set -u
TAG=${1:-n8}
mkdir -p gpurun_out
for N in ${NS:-8 4}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
      bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/${TAG}_bench_n$N.err | grep '^{"metric"' > gpurun_out/${TAG}_bench_n$N.json
  echo "bench N=$N rc=$?"; tail -c 400 gpurun_out/${TAG}_bench_n$N.err; head -c 1200 gpurun_out/${TAG}_bench_n$N.json; echo
done
