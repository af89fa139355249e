mkdir -p gpurun_out
rm -f gpurun_out/s5o_*
export GLC_MODEL_CACHE=/tmp/glc_models
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s5o_$name.json 2> gpurun_out/s5o_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/s5o_$name.json"))
print("$name", round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['kernels']['attention']['ms_per_step'], d['kernels']['residual_ln']['ms_per_step'])
PY
}
run all1 GLC_X=1
run v1 GLC_ATTN_SWAP=0 GLC_ATTN_OTMEM=0 GLC_ATTN_POLY=0
run nopoly GLC_ATTN_POLY=0
run all1b GLC_X=1
run v1b GLC_ATTN_SWAP=0 GLC_ATTN_OTMEM=0 GLC_ATTN_POLY=0
run stream GLC_ATTN=stream
