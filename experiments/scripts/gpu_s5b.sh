mkdir -p gpurun_out
rm -f gpurun_out/s5b_*
for g in 2 4; do
  echo "== G=$g" >> gpurun_out/s5b_attn.log
  GLC_ATTN_G=$g GLC_ATTN=shift timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s5b_attn.log 2>&1
  GLC_ATTN_G=$g GLC_ATTN=shift GLC_ATTN_TRACE=gpurun_out/s5b_trace_g$g.txt timeout 300 python scripts/bench_attn.py 64 512 12 1 >> gpurun_out/s5b_attn.log 2>&1
done
GLC_ATTN_G=4 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "shift" > gpurun_out/s5b_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s5b_kernels.log
tail -n 3 gpurun_out/s5b_kernels.log
grep -v parity gpurun_out/s5b_attn.log; cat gpurun_out/s5b_trace_g2.txt; cat gpurun_out/s5b_trace_g4.txt
