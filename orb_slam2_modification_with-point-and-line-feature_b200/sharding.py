"""Frame-wise sharding for batched offline extraction (BASELINE.json config 5).

Frames are independent units (ORBextractor::operator() keeps no state across calls), so rank r of G takes the
contiguous range [r*F/G, (r+1)*F/G).  There is no data-path collective: the only cross-rank step is the final gather
of the (variable-length) results — here `gather_results`, built on torch.distributed (NCCL on the GPUs, gloo in the
CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_frames: int, world: int, rank: int):
    """Contiguous [begin, end) of rank `rank`; ranges tile [0, n_frames) exactly, sizes differ by at most 1."""
    assert 0 <= rank < world and n_frames >= 0
    return (rank * n_frames) // world, ((rank + 1) * n_frames) // world


def frame_checksum(kps, desc) -> int:
    """Order-sensitive 63-bit checksum of one frame's keypoints and descriptors (FNV-1a over the raw bytes)."""
    h = 1469598103934665603
    data = np.ascontiguousarray(kps).tobytes() + np.ascontiguousarray(desc).tobytes()
    a = np.frombuffer(data, np.uint8)
    # vectorised chunked FNV: fold 8-byte words (exactness matters only for equality between runs)
    pad = (-len(a)) % 8
    w = np.frombuffer(data + b"\0" * pad, np.uint64)
    for x in (int(w.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(w)), len(a)):
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    mix = int((w * (np.arange(len(w), dtype=np.uint64) | np.uint64(1))).sum(dtype=np.uint64))
    h = ((h ^ mix) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h >> 1


def gather_results(local_counts, local_checksums, n_frames: int, world: int, rank: int, dist, device=None):
    """Gathers per-frame (count, checksum) of every rank's shard to rank 0 in frame order.  Returns two int64 arrays of
    length n_frames on rank 0 and None elsewhere."""
    import torch
    sizes = [shard_range(n_frames, world, r)[1] - shard_range(n_frames, world, r)[0] for r in range(world)]
    mx = max(sizes) if sizes else 0
    buf = torch.zeros((mx, 2), dtype=torch.int64, device=device)
    n = len(local_counts)
    if n:
        buf[:n, 0] = torch.as_tensor(np.asarray(local_counts, np.int64), device=device)
        buf[:n, 1] = torch.as_tensor(np.asarray(local_checksums, np.int64), device=device)
    if world == 1:
        out = [buf]
    else:
        out = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    counts = np.concatenate([out[r][:sizes[r], 0].cpu().numpy() for r in range(world)])
    sums = np.concatenate([out[r][:sizes[r], 1].cpu().numpy() for r in range(world)])
    return counts, sums
