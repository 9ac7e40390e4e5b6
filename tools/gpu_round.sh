#!/bin/bash
# One GPU-box visit: GPU test suite, then the default bench line.  Outputs under gpurun_out/.
set -u
TAG=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/${TAG}_bench.err
head -c 6000 gpurun_out/${TAG}_bench.json
