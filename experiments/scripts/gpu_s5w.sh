mkdir -p gpurun_out
rm -f gpurun_out/s5w_*
GLC_ATTN_GONCE=1 GLC_ATTN_POLY=3 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "shift and not stream" > gpurun_out/s5w_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/s5w_kernels.log
tail -n 4 gpurun_out/s5w_kernels.log
for cfg in "4 0" "3 0" "2 0" "0 0" "4 1" "3 1" "4 0"; do
  set -- $cfg
  echo "== POLY=$1 GONCE=$2" >> gpurun_out/s5w_attn.log
  GLC_ATTN_POLY=$1 GLC_ATTN_GONCE=$2 GLC_ATTN=shift timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s5w_attn.log 2>&1
done
grep -v "mode\|parity" gpurun_out/s5w_attn.log
