mkdir -p gpurun_out
rm -f gpurun_out/s5s_*
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "shift and not stream" > gpurun_out/s5s_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s5s_kernels.log
tail -n 12 gpurun_out/s5s_kernels.log
for h in 1 0 1 0; do
  echo "== C16=$h" >> gpurun_out/s5s_attn.log
  GLC_ATTN_C16=$h GLC_ATTN=shift timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s5s_attn.log 2>&1
done
grep -v "mode" gpurun_out/s5s_attn.log
