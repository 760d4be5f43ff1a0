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


def pack_features(kps, desc, n, kls, ldesc, lco, ln):
    """Compacts the padded per-frame outputs of a shard (torch tensors: kps (F, cap, 7) f32, desc (F, cap, 32) u8, n (F) i32, kls
    (F, L, 17) f32, ldesc (F, L, 32) u8, lco (F, L, 3) f64, ln (F) i32) into what travels in the final gather: the two count vectors
    and the five row lists without padding.  Works on any device."""
    import torch
    cap, L = kps.shape[1], kls.shape[1]

    def rows(cnt, per_frame):
        # flat row indices of the filled rows, frame after frame (what a boolean mask over (F, per_frame) would select, without the mask)
        c = cnt.to(torch.int64).clamp(0, per_frame)
        total = int(c.sum())
        first = torch.arange(len(c), device=c.device, dtype=torch.int64) * per_frame - (torch.cumsum(c, 0) - c)
        return torch.repeat_interleave(first, c, output_size=total) + torch.arange(total, device=c.device, dtype=torch.int64)

    pi, li = rows(n, cap), rows(ln, L)
    F = kps.shape[0]
    return [n.to(torch.int32).contiguous(), ln.to(torch.int32).contiguous(), kps.reshape(F * cap, -1).index_select(0, pi),
            desc.reshape(F * cap, -1).index_select(0, pi), kls.reshape(F * L, -1).index_select(0, li),
            ldesc.reshape(F * L, -1).index_select(0, li), lco.reshape(F * L, -1).index_select(0, li)]


def gather_features(parts, world: int, rank: int, dist, device=None):
    """The final gather of a frame-wise sharded extraction: every rank's packed features (pack_features) go to rank 0,
    which returns them concatenated in rank order == frame order (shards are contiguous frame ranges), and None elsewhere.
    Variable-length: the byte counts are gathered first, then every rank sends ONE buffer (NCCL send / recv on GPUs, gloo on CPU)."""
    import torch

    def pad16(nbytes: int) -> int:
        return (nbytes + 15) & ~15

    if world == 1:  # nothing travels
        return parts, int(sum(t.numel() * t.element_size() for t in parts))
    # every part starts at a multiple of 16 bytes inside the buffer, so the receiver can view it with the part's own dtype
    raw = [t.reshape(-1).view(torch.uint8) for t in parts]
    chunks = []
    for b in raw:
        chunks.append(b)
        extra = pad16(b.numel()) - b.numel()
        if extra:
            chunks.append(torch.zeros(extra, dtype=torch.uint8, device=b.device))
    flat = torch.cat(chunks) if chunks else torch.zeros(0, dtype=torch.uint8, device=device)
    sizes = torch.tensor([b.numel() for b in raw], dtype=torch.int64, device=flat.device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)] if rank == 0 else None
    dist.gather(sizes, all_sizes, dst=0)
    if rank != 0:
        dist.send(flat, dst=0)
        return None, int(sizes.sum())
    all_sizes = [[int(v) for v in a.cpu().tolist()] for a in all_sizes]
    bufs = [flat]
    for r in range(1, world):
        b = torch.empty(sum(pad16(v) for v in all_sizes[r]), dtype=torch.uint8, device=flat.device)
        dist.recv(b, src=r)
        bufs.append(b)
    out = []
    for k, proto in enumerate(parts):
        pieces = []
        for r in range(world):
            off = sum(pad16(v) for v in all_sizes[r][:k])
            pieces.append(bufs[r][off:off + all_sizes[r][k]].view(proto.dtype))
        cat = torch.cat(pieces)
        out.append(cat.reshape((-1,) + tuple(proto.shape[1:])) if proto.dim() > 1 else cat)
    return out, int(sum(sum(a) for a in all_sizes))


def features_checksum(parts) -> int:
    """Order-sensitive checksum of gathered features that does not depend on how the frames were sharded."""
    import torch
    h = 0
    for k, t in enumerate(parts):
        b = t.reshape(-1).view(torch.uint8).to(torch.int64)
        w = (torch.arange(b.numel(), device=b.device, dtype=torch.int64) % 65521) + 1
        h = (h * 1000003 + int((b * w).sum().item()) + 7919 * k + int(b.numel())) % ((1 << 61) - 1)
    return h
