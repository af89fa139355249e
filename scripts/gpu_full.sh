# full check on a GPU box: all -m gpu tests, smoke, bench (both arms).  scripts/gpu_profiles.sh adds the ncu evidence,
# scripts/gpu_configs.sh the other BASELINE configs.  Run through scripts/gpurun_retry.sh <timeout> scripts/gpu_full.sh [gpus]
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/f_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?" >> gpurun_out/f_bench.err
tail -n 4 gpurun_out/f_tests.log gpurun_out/f_smoke.log gpurun_out/f_bench.err
cat gpurun_out/f_bench.json gpurun_out/f_bench_ref.json
