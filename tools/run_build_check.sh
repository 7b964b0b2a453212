#!/bin/bash
# GPU check of the device-side tree build: parity tests, then the update latency comparison.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tree_build.py -x -q 2>&1 | tail -15
timeout 600 python tools/update_latency.py 2>&1 | tail -8
