mkdir -p gpurun_out
rm -f gpurun_out/s6f_*
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "shift and not stream" > gpurun_out/s6f_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s6f_kernels.log
tail -n 4 gpurun_out/s6f_kernels.log
for cfg in "GLC_X=1" "GLC_ATTN_C16=1 GLC_ATTN_POLY=3" "GLC_X=1" "GLC_ATTN_G=4"; do
  echo "== $cfg" >> gpurun_out/s6f_attn.log
  env $cfg GLC_ATTN=shift timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s6f_attn.log 2>&1
done
grep -v "mode" gpurun_out/s6f_attn.log
