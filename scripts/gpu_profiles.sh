# bench + reference arm + ncu evidence for the current build (outputs under gpurun_out/p_*); afterwards, here:
#   python tools/ncu_summary.py launches gpurun_out/p_launches.csv > profiles/rN_launches_summary.md
#   python tools/ncu_summary.py report gpurun_out/p_gemm.ncu-rep gpurun_out/p_attn.ncu-rep gpurun_out/p_ln.ncu-rep > profiles/rN_kernels_ncu.md
#   python tools/ncu_summary.py traffic gemm=gpurun_out/p_gemm.ncu-rep attention=gpurun_out/p_attn.ncu-rep residual_ln=gpurun_out/p_ln.ncu-rep > profiles/ncu_traffic.json
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; echo "bench rc=$?" >> gpurun_out/p_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/p_bench_ref.json 2> gpurun_out/p_bench_ref.err
NCU_BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/p_launches.csv $NCU_BENCH > gpurun_out/p_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_2cta -s 48 -c 4 -o gpurun_out/p_gemm -f $NCU_BENCH > gpurun_out/p_ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_persist -s 14 -c 1 -o gpurun_out/p_attn -f $NCU_BENCH > gpurun_out/p_ncu_attn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:residual_ln -s 28 -c 2 -o gpurun_out/p_ln -f $NCU_BENCH > gpurun_out/p_ncu_ln.log 2>&1
tail -n 2 gpurun_out/p_bench.err; cat gpurun_out/p_bench.json gpurun_out/p_bench_ref.json
ls -la gpurun_out/ | grep " p_"
