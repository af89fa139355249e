# new attention kernel: parity, micro-bench (new vs legacy), then e2e parity + bench
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "toeplitz" > gpurun_out/s4b_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s4b_kernels.log
tail -n 30 gpurun_out/s4b_kernels.log
if grep -q "rc=0" gpurun_out/s4b_kernels.log; then
  timeout 300 python scripts/bench_attn.py 64 512 12 20 > gpurun_out/s4b_attn.log 2>&1
  timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s4b_attn.log 2>&1
  timeout 300 python scripts/bench_attn.py 16 1024 16 10 >> gpurun_out/s4b_attn.log 2>&1
  timeout 300 python scripts/bench_attn.py 16 1024 16 10 >> gpurun_out/s4b_attn.log 2>&1
  cat gpurun_out/s4b_attn.log
  timeout 900 python -m pytest tests/test_gpu_e2e.py -x -q > gpurun_out/s4b_e2e.log 2>&1; echo "rc=$?" >> gpurun_out/s4b_e2e.log
  tail -n 5 gpurun_out/s4b_e2e.log
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s4b_bench.json 2> gpurun_out/s4b_bench.err; echo "bench rc=$?" >> gpurun_out/s4b_bench.err
  tail -n 3 gpurun_out/s4b_bench.err; cat gpurun_out/s4b_bench.json
fi
