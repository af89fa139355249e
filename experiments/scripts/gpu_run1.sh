# first-parity run: every GPU test group in its own process (a device trap poisons the context)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r1_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "gemm" 2>&1 | tail -150 > gpurun_out/r1_gemm.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "embed or residual or mask_prep or head" 2>&1 | tail -80 > gpurun_out/r1_ew.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "attention_naive" 2>&1 | tail -100 > gpurun_out/r1_attn_naive.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "attention_fused" 2>&1 | tail -250 > gpurun_out/r1_attn.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -s -k "not arch_parity and not base_arch" > gpurun_out/r1_e2e.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_e2e.py -q -s -k "arch_parity or base_arch" > gpurun_out/r1_e2e_big.log 2>&1
for f in gpurun_out/r1_*.log; do echo "== $f"; tail -n 3 $f; done
