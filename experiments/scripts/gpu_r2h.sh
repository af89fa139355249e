#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "rmsnorm or rope or swiglu or flash128" > gpurun_out/h_kernels.log 2>&1
echo "rc=$?" >> gpurun_out/h_kernels.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -s -k "qwen2_backbone_mini" > gpurun_out/h_e2e.log 2>&1
echo "rc=$?" >> gpurun_out/h_e2e.log
grep -v "^$" gpurun_out/h_kernels.log | tail -40; grep -E "qwen|passed|failed|Error|rc=" gpurun_out/h_e2e.log | tail -20
