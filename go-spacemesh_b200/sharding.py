"""Index-range sharding and the VRF-nonce min-reduction across ranks.

Labels are independent, so rank r of W takes a contiguous slice of the index range and no label bytes ever
cross ranks; the only exchange is one 48-byte candidate per rank (found, index, label32), gathered with
`torch.distributed.all_gather` (NCCL on GPUs, gloo in the CPU tests) and reduced locally with the same
order the kernels use: smallest label32 (big-endian), lowest index on ties.
"""
from __future__ import annotations

import numpy as np


def shard_range(start: int, count: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous shard of [start, start+count) for `rank`: same split as b200post_labels_range_multi."""
    per = (count + world - 1) // world
    off = min(per * rank, count)
    return start + off, min(per, count - off)


def encode_vrf(vrf) -> np.ndarray:
    """(index, label32) or None -> 6 x int64: [found, index, 4 big-endian label words]."""
    rec = np.zeros(6, dtype=np.int64)
    if vrf is not None:
        rec[0] = 1
        rec[1] = np.uint64(vrf[0]).astype(np.int64)
        rec[2:6] = np.frombuffer(vrf[1], dtype=">u8").astype(np.uint64).view(np.int64)
    return rec


def decode_vrf(rec: np.ndarray):
    if not rec[0]:
        return None
    label = np.asarray(rec[2:6]).astype(np.int64).view(np.uint64).astype(">u8").tobytes()
    return int(np.uint64(rec[1])), label


def reduce_vrf(records) -> tuple[int, bytes] | None:
    """Lexicographic min over (label32, index) of the found candidates."""
    best = None
    for rec in records:
        cand = decode_vrf(np.asarray(rec))
        if cand is not None and (best is None or (cand[1], cand[0]) < (best[1], best[0])):
            best = cand
    return best


def allgather_vrf(vrf, device=None):
    """All ranks contribute their candidate; every rank returns the global minimum."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return vrf
    mine = torch.from_numpy(encode_vrf(vrf)).to(device if device is not None else "cpu")
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return reduce_vrf([t.cpu().numpy() for t in out])
