"""N>1 host logic of the hot path on CPU: world_size-2 gloo processes.  The per-row
compute is replaced by the CPU oracle (test infrastructure) so that the sharding, the global class
count and the host gather are checked without a GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def _sharding():
    pkg = graft.load_package()
    import importlib
    return importlib.import_module(pkg.__name__ + ".sharding")


def test_shard_rows_partition():
    S = _sharding()
    for B in (0, 1, 7, 8, 64, 4096, 4099):
        for G in (1, 2, 3, 4, 8):
            spans = [S.shard_rows(B, G, r) for r in range(G)]
            covered = [i for lo, hi in spans for i in range(lo, hi)]
            assert covered == list(range(B))                      # contiguous, disjoint, complete, ordered
            assert max(hi - lo for lo, hi in spans) <= (B + G - 1) // G
    with pytest.raises(ValueError):
        S.shard_rows(8, 2, 2)


def test_count_classes_is_row_max():
    S = _sharding()
    ids = np.array([[1, 9, 9, 5], [9, 9, 9, 2], [1, 2, 3, 4]])
    assert S.count_classes(ids, 9) == 3
    assert S.count_classes(ids[2:], 9) == 0
    assert S.count_classes(np.zeros((0, 4), np.int64), 9) == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S = _sharding()
        orc = graft.load_oracle()
        cfg = orc.make_config("tiny")
        w = orc.init_weights(cfg, 0)
        # 7 rows (odd: shards of 4 and 3), mixed label counts: the widest row sits in rank 1's shard, so rank 0
        # only produces the reference's output width if C is agreed globally
        ids, mask = orc.synth_inputs(cfg, 7, 96, [2, 1, 2, 1, 2, 4, 1], seed=77, ragged=True)
        calls = []

        def run_rows(i, m, C):
            calls.append((i.shape[0], C))
            out = orc.forward_restated(w, cfg, torch.from_numpy(i), torch.from_numpy(m)).numpy()
            full = np.zeros((i.shape[0], C), np.float32)      # zero-padded classes still score:
            full[:, : out.shape[1]] = out                     # only the columns this shard has are compared below
            return full

        got = S.run_sharded(run_rows, ids.numpy(), mask.numpy(), cfg.class_token_index)
        lo, hi = S.shard_rows(7, world, rank)
        local_c = S.global_num_classes(ids.numpy()[lo:hi], cfg.class_token_index)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), got=got, calls=np.array(calls), local_c=local_c)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_run_matches_single_process(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    orc = graft.load_oracle()
    cfg = orc.make_config("tiny")
    w = orc.init_weights(cfg, 0)
    ids, mask = orc.synth_inputs(cfg, 7, 96, [2, 1, 2, 1, 2, 4, 1], seed=77, ragged=True)
    want = orc.forward_restated(w, cfg, ids, mask).numpy()          # [7, 4]
    res = [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]
    assert np.array_equal(res[0]["got"], res[1]["got"])              # every rank holds the gathered result
    got = res[0]["got"]
    assert got.shape == want.shape == (7, 4)
    # rows 0-3 were computed by rank 0 (which saw at most 2 labels locally), rows 4-6 by rank 1
    assert res[0]["calls"].tolist() == [[4, 4]] and res[1]["calls"].tolist() == [[3, 4]]
    assert int(res[0]["local_c"]) == 4 and int(res[1]["local_c"]) == 4    # MAX all_reduce of the class count
    counts = [2, 1, 2, 1, 2, 4, 1]
    for b in range(7):
        # each shard's oracle call only had its own widest row's width: compare the classes that row has
        width = max(counts[:4]) if b < 4 else max(counts[4:])
        np.testing.assert_allclose(got[b, :width], want[b, :width], atol=1e-5)
