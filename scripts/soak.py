"""Soak / fuzz run on a GPU box: random ragged requests through the packed and the padded layout (must agree bit for bit),
concurrent callers, repeated long steps.  Exits non-zero on any mismatch; every kernel wait is bounded, so a protocol bug
shows up as a trap, not a hang.  Usage: python scripts/soak.py [iterations]"""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
import bench
from tools import synth_model as SM

pkg = graft.load_package()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(2024)
fails = 0
for arch in ("base", "qwen-mini"):
    path = bench.model_path(arch)
    cfg = SM.make_config(arch)
    SM.make_model_file(arch, path, seed=0)
    os.environ["GLC_VARLEN"] = "0"
    s_pad = pkg.Session(path)
    os.environ.pop("GLC_VARLEN")
    s_pk = pkg.Session(path)
    t0 = time.time()
    for it in range(iters):
        B = int(rng.integers(16, 65)); S = int(rng.choice([256, 300, 384, 512, 640, 1000, 1024]))
        labels = [int(x) for x in rng.integers(1, 12, size=B)]
        ids, mask = SM.synth_inputs(cfg, B, S, labels, seed=int(rng.integers(1 << 30)), ragged=True, min_frac=float(rng.choice([0.15, 0.3, 0.6, 0.95])))
        a = s_pad.run_inference(ids.numpy(), mask.numpy())
        b = s_pk.run_inference(ids.numpy(), mask.numpy())
        if not np.array_equal(a, b) or not np.isfinite(a).all():
            fails += 1
            print(f"MISMATCH {arch} it={it} B={B} S={S}: max|d|={np.abs(a - b).max():.3e}", flush=True)
    st = s_pk.packed_stats()
    print(f"{arch}: {iters} random ragged requests, packed launches {st[0]}, rows {st[1]} of {st[2]}, {time.time() - t0:.1f}s, mismatches so far {fails}", flush=True)
    # concurrent callers with different shapes on the packed session
    errs = []
    def worker(k):
        try:
            r = np.random.default_rng(k)
            for _ in range(6):
                B = int(r.integers(4, 40)); S = int(r.choice([128, 256, 512]))
                ids, mask = SM.synth_inputs(cfg, B, S, 4, seed=int(r.integers(1 << 30)), ragged=True, min_frac=0.3)
                x = s_pk.run_inference(ids.numpy(), mask.numpy())
                y = s_pad.run_inference(ids.numpy(), mask.numpy())
                if np.abs(x - y).max() > 5e-3:
                    errs.append((k, B, S, float(np.abs(x - y).max())))
        except Exception as e:   # noqa: BLE001
            errs.append((k, repr(e)))
    th = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    [t.start() for t in th]; [t.join() for t in th]
    print(f"{arch}: 8 concurrent callers x 6 requests: {'ok' if not errs else errs[:3]}", flush=True)
    fails += len(errs)
    s_pad.close(); s_pk.close()
print("SOAK", "FAILED" if fails else "OK")
sys.exit(1 if fails else 0)
