mkdir -p gpurun_out
rm -f gpurun_out/s6d_*
for w in 0 1 0 1; do
  echo "== WHATIF=$w" >> gpurun_out/s6d_attn.log
  GLC_ATTN_WHATIF=$w GLC_ATTN_FLAGS=1 GLC_ATTN=shift timeout 300 python scripts/bench_attn.py 64 512 12 20 >> gpurun_out/s6d_attn.log 2>&1
done
grep -v "mode\|parity" gpurun_out/s6d_attn.log
export GLC_MODEL_CACHE=/tmp/glc_models
for w in 0 1; do
GLC_ATTN_WHATIF=$w timeout 600 python bench.py --gpus 1 --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/s6d_bench$w.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/s6d_bench$w.json')); print('whatif $w', round(d['value'],1), d['ms_per_step'], d['clocks']['sm_mhz'], d['kernels']['attention']['ms_per_step'])"
done
