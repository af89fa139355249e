"""In-process sharded C4-shaped run (512 texts x 1024 x 100 labels per device) with host-side timing of its parts."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
import bench
from tools import synth_model as SM
pkg = graft.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
path = bench.model_path("base")
cfg = SM.make_config("base")
SM.make_model_file("base", path, seed=0)
sess = pkg.Session(path, devices=list(range(n)))
i4, m4 = SM.synth_inputs(cfg, 64, 1024, 100, seed=77)
reps = 512 * n // 64
ids = np.ascontiguousarray(np.tile(i4.numpy(), (reps, 1)))
mask = np.ascontiguousarray(np.tile(m4.numpy(), (reps, 1)))
sess.run_inference(ids[:16 * n], mask[:16 * n])
out = np.empty((512 * n, 100), dtype=np.float32)
t0 = time.perf_counter(); C = sess.num_classes(ids); t_nc = time.perf_counter() - t0
ts = []
for _ in range(4):
    t0 = time.perf_counter()
    sess.run_inference(ids, mask, out=out)
    ts.append(time.perf_counter() - t0)
print(f"devices {n}: num_classes scan {t_nc * 1e3:.2f} ms; runs {[round(t * 1e3, 1) for t in ts]} ms; best {512 * n / min(ts):.0f} texts/s = {512 / min(ts):.0f} per device")
