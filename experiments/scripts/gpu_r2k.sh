#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_batch.py base > gpurun_out/k_prof.txt 2>&1
GLC_NO_PDL=1 python scripts/prof_batch.py base > gpurun_out/k_prof_nopdl.txt 2>&1
cat gpurun_out/k_prof.txt; tail -12 gpurun_out/k_prof_nopdl.txt
