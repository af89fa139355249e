"""Batch sharding of the hot path over GPUs / ranks.

Rows of a batch are independent — the reference's OpenMP loop (main.c:141-150) hands each batch of
BATCH_SIZE texts to one `Run`, and inside a `Run` no op mixes rows except `C = max_b count(<<LABEL>>)`,
which only fixes the output WIDTH.  So the multi-GPU plan is: contiguous row shards, a full weight
replica per GPU, NO data-path collective; the per-rank [rows, C] fp32 logits are gathered on the host
(one process, `Model::run` in csrc/engine.cu) or with one all_gather of a few KB (one process per
GPU under torchrun: this module).  The only value that must be agreed globally is C.

Pure host logic (numpy + optional torch.distributed); no CUDA here, so it is tested on CPU with the
gloo backend (tests/test_sharding.py).
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np


def shard_rows(num_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of the contiguous row shard of `rank`: ceil(B/G) rows per shard, as
    Model::run (csrc/engine.cu) splits one large Run; trailing ranks may get an empty shard."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    per = (num_rows + world - 1) // world
    start = min(rank * per, num_rows)
    return start, min(start + per, num_rows)


def count_classes(input_ids: np.ndarray, class_token: int) -> int:
    """Output width of the reference graph: max over rows of the number of <<LABEL>> tokens
    (the `Equal(input_ids, class_token_index)` + `ReduceSum` + `ReduceMax` of the traced graph)."""
    if input_ids.size == 0:
        return 0
    return int((input_ids == class_token).sum(axis=1).max())


def global_num_classes(local_ids: np.ndarray, class_token: int, group=None) -> int:
    """C over ALL shards (MAX all_reduce of one int64) so every rank produces the reference's width."""
    c = count_classes(local_ids, class_token)
    try:
        import torch
        import torch.distributed as dist
    except Exception:   # noqa: BLE001
        return c
    if not (dist.is_available() and dist.is_initialized()):
        return c
    t = torch.tensor([c], dtype=torch.int64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def run_sharded(run_rows: Callable[[np.ndarray, np.ndarray, int], np.ndarray], input_ids: np.ndarray,
                attention_mask: np.ndarray, class_token: int, group=None) -> np.ndarray:
    """One logical `Run` over a batch that every rank holds in full: each rank computes its row shard
    with `run_rows(ids, mask, C) -> [rows, C] fp32` (Session.run_inference on a GPU; any callable in
    tests) and the shards are gathered into the full [B, C] matrix on every rank.

    Host gather = all_gather of the padded shards (≤ 1.6 MB at 4096 texts x 100 labels); there is no
    collective inside `run_rows`."""
    import torch
    import torch.distributed as dist

    ids = np.ascontiguousarray(input_ids, dtype=np.int64)
    mask = np.ascontiguousarray(attention_mask, dtype=np.int64)
    B = ids.shape[0]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    C = count_classes(ids, class_token)          # every rank sees the whole batch here
    lo, hi = shard_rows(B, world, rank)
    local = run_rows(ids[lo:hi], mask[lo:hi], C) if hi > lo else np.zeros((0, C), np.float32)
    local = np.asarray(local, dtype=np.float32).reshape(hi - lo, C)
    if world == 1:
        return local
    per = (B + world - 1) // world
    pad = np.zeros((per, C), np.float32)
    pad[: hi - lo] = local
    mine = torch.from_numpy(pad)
    parts: List[torch.Tensor] = [torch.empty_like(mine) for _ in range(world)]
    if dist.get_backend(group) == "nccl":
        mine = mine.cuda()
        parts = [p.cuda() for p in parts]
    dist.all_gather(parts, mine, group=group)
    full = torch.cat([p.cpu() for p in parts], dim=0)[:B]
    return full.numpy()
