# full check: all -m gpu tests, smoke, bench (both arms)
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/f_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/f_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?" >> gpurun_out/f_bench.err
tail -n 4 gpurun_out/f_tests.log gpurun_out/f_smoke.log gpurun_out/f_bench.err
cat gpurun_out/f_bench.json
