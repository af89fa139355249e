export GLC_MODEL_CACHE=/tmp/glc_models
timeout 900 python -m pytest tests/test_gpu_e2e.py -x -q -s -k "full_size_properties" 2>&1 | tail -8
