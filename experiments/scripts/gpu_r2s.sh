#!/bin/bash
# 2 GPUs: full -m gpu suite (incl. the two-device sharding test), in-process probe
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/s_tests.log
tail -6 gpurun_out/s_tests.log
python scripts/inproc_probe.py 2 2>&1 | grep devices
