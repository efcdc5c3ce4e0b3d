"""(clip x head) partitioning of the memory op across the B200s of one box.

Time is sequential but every (clip, head) chain is independent, so ranks own
disjoint blocks of clips -- all heads of a clip stay together so q/k/v slices remain contiguous --
and the hot path needs NO collective.  The only exchange is the optional readout gather after the
op (``gather_readout``: one NCCL all_gather over NVLink/NVSwitch), reported separately by bench.py.
Heads are split instead when there are fewer clips than ranks.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def chain_partition(B: int, H: int, world: int, rank: int) -> Tuple[slice, slice]:
    """(clip slice, head slice) owned by ``rank``.  Balanced to within one clip (or one head)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    if B >= world:
        base, rem = divmod(B, world)
        b0 = rank * base + min(rank, rem)
        return slice(b0, b0 + base + (1 if rank < rem else 0)), slice(0, H)
    # fewer clips than ranks: ranks_per_clip ranks share one clip and split its heads
    if world % B != 0 or H % (world // B) != 0:
        raise ValueError(f"cannot split B={B}, H={H} over {world} ranks evenly")
    rpc = world // B
    hb = H // rpc
    return slice(rank // rpc, rank // rpc + 1), slice((rank % rpc) * hb, (rank % rpc + 1) * hb)


def packed_partition(cu_seqlens, world: int, rank: int) -> Tuple[slice, slice]:
    """(clip slice, token slice) of packed variable-length clips owned by ``rank``: contiguous blocks of whole clips whose
    token counts are as even as a prefix split allows (a clip is never cut: its chains are sequential in time).  The rank
    runs ``gdr_lkva_varlen`` on ``q[:, tokens]`` with ``cu_seqlens[clips.start : clips.stop + 1] - cu_seqlens[clips.start]``;
    no collective is needed.  ``cu_seqlens``: host sequence of N + 1 offsets."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    cu = [int(x) for x in cu_seqlens]
    n, total = len(cu) - 1, cu[-1]
    bounds = [0]
    for r in range(1, world):                    # clip boundary closest to r / world of the tokens, monotone
        target = total * r / world
        j = min(range(bounds[-1], n + 1), key=lambda i: (abs(cu[i] - target), i))
        bounds.append(j)
    bounds.append(n)
    lo, hi = bounds[rank], bounds[rank + 1]
    return slice(lo, hi), slice(cu[lo], cu[hi])


def gather_readout(o_local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather the per-rank readout [B_r,T,H,V] along the clip dimension (equal B_r per rank)."""
    world = dist.get_world_size(group)
    out = o_local.new_empty((world * o_local.shape[0],) + tuple(o_local.shape[1:]))
    dist.all_gather_into_tensor(out, o_local.contiguous(), group=group)
    return out
