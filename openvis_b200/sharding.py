"""Clip sharding across ranks and the end-of-run result gather.

Mirrors the reference's evaluation pattern: `InferenceSampler` gives every rank a contiguous block of videos
(openvis/data/build.py:238-247) and the per-video results are gathered on rank 0 once, at the end
(openvis/data/evals/ytvis_eval.py:117-128, there via pickles over gloo).  Here the results are fixed-shape tensors
and the gather is one `all_gather` (NCCL over NVLink on GPUs; gloo in the CPU tests).  Nothing on the decoding path
communicates.
"""
from typing import List

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous block of `n_items` for `rank` (same split as detectron2's InferenceSampler: the first
    n % world ranks get one extra item)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return range(begin, begin + base + (1 if rank < rem else 0))


def gather_clip_results(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """local: [n_local, ...] results of this rank's block, in block order.  Returns [n_items, ...] in global clip order
    on every rank.  Blocks may differ in length by one; shorter blocks are padded for the collective."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if n_items % world == 0 and local.shape[0] == n_items // world and local.is_contiguous():
        # equal blocks: rank order IS global clip order, so the collective writes the result in place -- no padding copy, no
        # list of parts, no concatenation (at 8 GPUs those copies were a third of the 5.9 ms gather of 2.7 GB of packed masks)
        out = torch.empty((n_items,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out.view(torch.uint8).reshape(-1), local.view(torch.uint8).reshape(-1), group=group)
        return out
    n_max = -(-n_items // world)
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    # the collective moves raw bytes, so every element type works on both back-ends (neither NCCL nor gloo has int16)
    raw = pad.reshape(n_max, -1).view(torch.uint8)
    parts: List[torch.Tensor] = [torch.empty_like(raw) for _ in range(world)]
    dist.all_gather(parts, raw, group=group)
    out = [parts[r].view(local.dtype).reshape(pad.shape)[: len(shard_range(n_items, r, world))] for r in range(world)]
    return torch.cat(out, dim=0)


def gather_clip_dict(local: dict, n_items: int, group=None) -> dict:
    """The fixed-shape per-clip results of one rank's block -- e.g. class scores [n, Q, K] fp32, query-matching indices
    [n, T, Q] (int64 as temporal.batch_video_match_via_embeds returns them, or narrowed to int16 by the caller), top-10 ids and bit-packed masks int32 -- gathered key by key
    into global clip order (SURVEY.md section 8 e).  One collective per key, issued once per run."""
    return {k: gather_clip_results(v, n_items, group) for k, v in local.items()}
