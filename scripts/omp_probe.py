"""The reference's calling pattern (main.c:141-150): T host threads x synchronous batch-8 Runs; texts/s and how the engine
grouped them.  Usage: python scripts/omp_probe.py [threads runs_per_thread]"""
import os, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
import bench
from tools import synth_model as SM
pkg = graft.load_package()
nthr, per = (list(map(int, sys.argv[1:3])) + [16, 12][len(sys.argv) - 1:])[:2]
path = bench.model_path("base"); cfg = SM.make_config("base"); SM.make_model_file("base", path, seed=0)
sess = pkg.Session(path)
S, NL = 512, 10
bufs = []
for t in range(nthr):
    i8, m8 = SM.synth_inputs(cfg, 8, S, NL, seed=5000 + t)
    bufs.append((i8.pin_memory(), m8.pin_memory(), torch.empty(8, NL).pin_memory()))
def worker(t, n):
    i8, m8, o8 = bufs[t]
    for _ in range(n):
        sess.run_pinned(i8.data_ptr(), m8.data_ptr(), 8, S, o8.data_ptr(), o8.numel())
def round_(n):
    th = [threading.Thread(target=worker, args=(t, n)) for t in range(nthr)]
    t0 = time.perf_counter(); [x.start() for x in th]; [x.join() for x in th]
    return time.perf_counter() - t0
round_(2)
for rep in range(3):
    g0, r0 = sess.coalesce_stats(); l0 = sess.launch_count()
    dt = round_(per)
    g1, r1 = sess.coalesce_stats()
    print(f"{nthr} threads x {per} runs: {nthr * per * 8 / dt:.0f} texts/s; merged launches {g1 - g0} serving {r1 - r0} of {nthr * per} requests")
