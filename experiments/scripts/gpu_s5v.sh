mkdir -p gpurun_out
rm -f gpurun_out/s5v_*
GLC_ATTN_G=4 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "shift and not stream" > gpurun_out/s5v_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s5v_kernels.log
tail -n 4 gpurun_out/s5v_kernels.log
for g in 4 2 4 2; do
  echo "== G=$g" >> gpurun_out/s5v_attn.log
  GLC_ATTN_G=$g GLC_ATTN=shift timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s5v_attn.log 2>&1
done
grep -v "mode" gpurun_out/s5v_attn.log
