#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/g_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/g_tests.log
tail -8 gpurun_out/g_tests.log
