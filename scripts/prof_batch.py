"""Per-kernel-family device time of one forward at a given batch (in-stream profiler), next to batch 64 / (64/B)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
import __graft_entry__ as graft
from tools import synth_model as SM
glc = graft.load_package()
arch = sys.argv[1] if len(sys.argv) > 1 else "base"
path = bench.model_path(arch)
cfg = SM.make_config(arch)
SM.make_model_file(arch, path, seed=0)
sess = glc.Session(path)
S, NL = 512, 10
out = {}
for B in (8, 64):
    ids, mask = SM.synth_inputs(cfg, B, S, NL, seed=1235)
    C = sess.num_classes(ids.numpy())
    d_ids = ids.cuda(); d_mask = mask.cuda()
    d_out = torch.empty(B, C, dtype=torch.float32, device="cuda")
    for _ in range(5): sess.run_device(d_ids.data_ptr(), d_mask.data_ptr(), B, S, C, d_out.data_ptr(), sync=False)
    torch.cuda.synchronize()
    sess.profile_enable(True); sess.profile_collect()
    n = 20
    for _ in range(n): sess.run_device(d_ids.data_ptr(), d_mask.data_ptr(), B, S, C, d_out.data_ptr(), sync=False)
    prof = sess.profile_collect(); sess.profile_enable(False)
    out[B] = {k: (v[0] / n * 1e3, v[1] / n) for k, v in prof.items()}
print(f"{'family':14s} {'B=8 us':>9s} {'n':>4s} {'B=64/8 us':>10s} {'ratio':>6s}")
t8 = t64 = 0
for k in out[8]:
    a, na = out[8][k]; b = out[64][k][0] / 8
    t8 += a; t64 += b
    print(f"{k:14s} {a:9.1f} {na:4.0f} {b:10.1f} {a / b if b else 0:6.2f}")
print(f"{'total':14s} {t8:9.1f}      {t64:10.1f} {t8 / t64:6.2f}")
