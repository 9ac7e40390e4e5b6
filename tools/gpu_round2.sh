#!/bin/bash
# Two-GPU box visit: the whole GPU suite (incl. tests/test_gpu_multi.py) and a 2-rank bench line.
set -u
TAG=${1:-run2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 3 --warmup 3 2> gpurun_out/${TAG}_bench_n2.err | grep "^{\"metric\"" > gpurun_out/${TAG}_bench_n2.json
echo "bench rc=$?"
tail -c 2000 gpurun_out/${TAG}_bench_n2.err
head -c 3000 gpurun_out/${TAG}_bench_n2.json
