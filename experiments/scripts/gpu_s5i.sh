# other BASELINE configs (parity-test cases, measured for orientation) and the batch-8 latency shape, kernel breakdowns
mkdir -p gpurun_out
rm -f gpurun_out/s5i_*
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 600 python bench.py --no-cpu-baseline --arch small --batch 8 --seq 512 --labels 4 --steps 20 --warmup 5 > gpurun_out/s5i_c1.json 2> gpurun_out/s5i_c1.err
timeout 600 python bench.py --no-cpu-baseline --arch base --batch 8 --seq 512 --labels 10 --steps 20 --warmup 5 > gpurun_out/s5i_b8.json 2> gpurun_out/s5i_b8.err
timeout 900 python bench.py --no-cpu-baseline --arch large --batch 128 --seq 1024 --labels 50 --steps 5 --warmup 3 > gpurun_out/s5i_c3.json 2> gpurun_out/s5i_c3.err
timeout 900 python bench.py --no-cpu-baseline --arch base --batch 512 --seq 1024 --labels 100 --steps 3 --warmup 3 > gpurun_out/s5i_c4.json 2> gpurun_out/s5i_c4.err
python - <<'PY'
import json
for n in ["c1","b8","c3","c4"]:
    try:
        d=json.load(open(f"gpurun_out/s5i_{n}.json"))
        print(n, d["config"]["workload"], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), "frac", round(d["config"]["whole_forward_frac_of_tensor_peak"],3))
        print("   ", {k:v["ms_per_step"] for k,v in d["kernels"].items()})
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/s5i_{n}.err").read()[-800:])
PY
