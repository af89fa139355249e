#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
python scripts/inproc_probe.py 1 > gpurun_out/r_inproc.txt 2>&1
python scripts/inproc_probe.py 8 >> gpurun_out/r_inproc.txt 2>&1
GLC_TIMING=1 python scripts/inproc_probe.py 8 > gpurun_out/r_inproc_timing.txt 2>&1
grep devices gpurun_out/r_inproc.txt; tail -30 gpurun_out/r_inproc_timing.txt
