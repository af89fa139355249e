mkdir -p gpurun_out
rm -f gpurun_out/s6e_*
for w in 0 2 4 8 14 15 0; do
  echo "== WHATIF=$w" >> gpurun_out/s6e_attn.log
  GLC_ATTN_WHATIF=$w GLC_ATTN_FLAGS=1 GLC_ATTN=shift timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s6e_attn.log 2>&1
done
grep -v "mode\|parity" gpurun_out/s6e_attn.log
