mkdir -p gpurun_out
rm -f gpurun_out/s5a_*
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "shift" > gpurun_out/s5a_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s5a_kernels.log
tail -n 30 gpurun_out/s5a_kernels.log
GLC_ATTN=shift timeout 300 python scripts/bench_attn.py 64 512 12 20 > gpurun_out/s5a_attn.log 2>&1
GLC_ATTN=gather timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s5a_attn.log 2>&1
cat gpurun_out/s5a_attn.log
if grep -q "rc=0" gpurun_out/s5a_kernels.log; then
  GLC_ATTN=shift timeout 900 python -m pytest tests/test_gpu_e2e.py -x -q > gpurun_out/s5a_e2e.log 2>&1; echo "rc=$?" >> gpurun_out/s5a_e2e.log
  tail -n 5 gpurun_out/s5a_e2e.log
  GLC_ATTN=shift timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s5a_bench_shift.json 2> gpurun_out/s5a_bench_shift.err
  cat gpurun_out/s5a_bench_shift.json
fi
