"""Multi-GPU plumbing: utterances are independent (the reference is batch-1,
/root/reference/src/inference.rs:89), so a batch shards one block of utterances per rank with
weights replicated, and the ONLY collective is the final gather of generated token ids
(SURVEY.md section 8e).  One process per GPU; torch.distributed is plumbing (NCCL over
NVLink on GPUs, gloo on CPU for the tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of utterance indices owned by ``rank`` (ceil split)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def gather_token_ids(local_ids: Sequence[Sequence[int]], n_items: int, max_new: int, device=None,
                     group=None) -> List[List[int]]:
    """All ranks receive the ids of all ``n_items`` utterances in global order.
    Payload: int32 [per, max_new + 1] per rank (length in column 0) -- <= 512 KB, latency-bound."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (n_items + world - 1) // world
    buf = torch.full((per, max_new + 1), -1, dtype=torch.int32)
    for i, ids in enumerate(local_ids):
        n = min(len(ids), max_new)
        buf[i, 0] = n
        if n:
            buf[i, 1:1 + n] = torch.tensor(list(ids[:n]), dtype=torch.int32)
    if device is not None:
        buf = buf.to(device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    res: List[List[int]] = []
    for r in range(world):
        lo, hi = shard_range(n_items, world, r)
        t = out[r].cpu()
        for i in range(hi - lo):
            n = int(t[i, 0])
            res.append(t[i, 1:1 + n].tolist())
    return res
