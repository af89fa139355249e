#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/p2_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/p2_tests.log
tail -5 gpurun_out/p2_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
