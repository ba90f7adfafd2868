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


class _DevView:
    """Zero-copy view of library-owned device memory for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape, typestr: str = "<i4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),   # torch rejects read-only views; IdsGather only reads
                                         "version": 3, "strides": None}


class IdsGather:
    """The gather of decoded ids for engines that keep them in HBM: ONE ``all_gather_into_tensor`` (NCCL over NVLink)
    reading the session's own device buffers (asrb_session_device_ids) and one D2H of the gathered block -- no host
    staging on the way in, no per-rank tensor list.  Buffers are allocated once.

        g = IdsGather(world, n_items, max_new, device)      # n_items = utterances of the whole job
        all_ids = g(engine)                                  # after engine.transcribe_ids(...)
    """

    def __init__(self, world: int, n_items: int, max_new: int, device, group=None):
        import torch
        self.world, self.n_items, self.max_new, self.group = world, n_items, max_new, group
        self.per = (n_items + world - 1) // world
        self.send = torch.full((self.per, max_new + 1), -1, dtype=torch.int32, device=device)
        self.recv = torch.empty((world * self.per, max_new + 1), dtype=torch.int32, device=device)
        self.host = torch.empty((world * self.per, max_new + 1), dtype=torch.int32).pin_memory() if device is not None and \
            getattr(device, "type", "cpu") == "cuda" else torch.empty((world * self.per, max_new + 1), dtype=torch.int32)

    def __call__(self, engine) -> List[List[int]]:
        import torch
        import torch.distributed as dist
        ids_ptr, lens_ptr, stride, batch = engine.device_ids()
        n = min(batch, self.per)
        ids = torch.as_tensor(_DevView(ids_ptr, (batch, stride)), device=self.send.device)
        lens = torch.as_tensor(_DevView(lens_ptr, (batch,)), device=self.send.device)
        w = min(stride, self.max_new)
        self.send[:n, 0] = lens[:n].clamp(max=w)
        self.send[:n, 1:1 + w] = ids[:n, :w]
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        self.host.copy_(self.recv, non_blocking=False)
        res: List[List[int]] = []
        for r in range(self.world):
            lo, hi = shard_range(self.n_items, self.world, r)
            for i in range(hi - lo):
                row = self.host[r * self.per + i]
                res.append(row[1:1 + int(row[0])].tolist())
        return res
