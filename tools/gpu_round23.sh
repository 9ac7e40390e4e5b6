#!/bin/bash
# the GPU suite against a library built with -DOVL_BOUNDS_CHECKS=1 (compute-sanitizer is closed on this pool), then against the default one
set -u
TAG=${1:-run23}
mkdir -p gpurun_out
OVL_B200_LIB=build/variants/libovl_bounds.so python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_bounds.log 2>&1
echo "pytest (bounds-checked library) rc=$?"; tail -3 gpurun_out/${TAG}_pytest_bounds.log
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest (default) rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
