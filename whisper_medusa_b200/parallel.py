"""Multi-GPU plumbing of the path: streams shard, weights are broadcast once.

The reference's inference is single-GPU, batch 1 (``model.py:1451``).  Independent audio streams
shard naturally: one process per GPU, every process holds a full weight replica and decodes its
own streams -- there is no data-path collective.  The only collective is the start-up broadcast
of the packed weight blob from the rank that read the checkpoint (NCCL over NVLink on GPUs; the
same code runs over gloo on CPU tensors in the tests).
"""
from __future__ import annotations

from typing import List, Optional

import torch


def stream_ids_for_rank(rank: int, world: int, n_streams: int) -> List[int]:
    """Round-robin assignment: stream s is decoded by rank ``s % world``."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_streams, world))


def broadcast_packed_weights(nbytes: int, src: int, blob: Optional[torch.Tensor], device: torch.device) -> torch.Tensor:
    """One broadcast of the packed uint8 blob.  ``blob`` is only needed on ``src`` (host tensor);
    returns the tensor on ``device`` on every rank (the engine adopts its pointer)."""
    import torch.distributed as dist

    out = torch.empty(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        if blob is None or blob.numel() != nbytes:
            raise ValueError("source rank must provide the packed blob")
        out.copy_(blob)
    dist.broadcast(out, src=src)
    return out


def gather_stream_results(local: List[List[int]], local_ids: List[int], n_streams: int) -> Optional[List[List[int]]]:
    """Collect per-stream token lists on rank 0 in stream order (host-side, tiny)."""
    import torch.distributed as dist

    world = dist.get_world_size()
    payload = list(zip(local_ids, local))
    gathered: List[Optional[list]] = [None] * world
    dist.all_gather_object(gathered, payload)
    if dist.get_rank() != 0:
        return None
    out: List[Optional[List[int]]] = [None] * n_streams
    for part in gathered:
        for sid, toks in part:
            out[sid] = toks
    if any(o is None for o in out):
        raise RuntimeError("a stream was not decoded by any rank")
    return out  # type: ignore[return-value]
