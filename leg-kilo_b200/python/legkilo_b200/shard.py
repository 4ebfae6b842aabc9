"""Multi-GPU partition of a batch of independent scans: contiguous batch/G scans per
rank, map replicated per GPU, no collective on the hot path. Host logic only."""
from __future__ import annotations

import numpy as np


def shard_range(n_scans: int, rank: int, world: int) -> tuple[int, int]:
    """Scans [lo, hi) owned by `rank`: contiguous, sizes differ by at most one, earlier ranks larger."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    base, rem = divmod(n_scans, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(rank: int, world: int, x, P, clk, pts, scan_offsets, bucket_times, scan_bucket_ptr=None, bucket_offsets=None):
    """Slice every per-scan array of a batch down to this rank's scans and rebase the offsets."""
    scan_offsets = np.asarray(scan_offsets, np.uint32)
    batch = len(scan_offsets) - 1
    lo, hi = shard_range(batch, rank, world)
    if scan_bucket_ptr is None:
        scan_bucket_ptr = np.arange(batch + 1, dtype=np.uint32)
        bucket_offsets = scan_offsets
    scan_bucket_ptr = np.asarray(scan_bucket_ptr, np.uint32)
    bucket_offsets = np.asarray(bucket_offsets, np.uint32)
    p0, p1 = int(scan_offsets[lo]), int(scan_offsets[hi])
    b0, b1 = int(scan_bucket_ptr[lo]), int(scan_bucket_ptr[hi])
    return dict(x=np.asarray(x)[lo:hi], P=np.asarray(P).reshape(batch, -1)[lo:hi], clk=np.asarray(clk)[lo:hi],
                pts=np.asarray(pts).reshape(-1, 4)[p0:p1], scan_offsets=(scan_offsets[lo:hi + 1] - p0).astype(np.uint32),
                scan_bucket_ptr=(scan_bucket_ptr[lo:hi + 1] - b0).astype(np.uint32),
                bucket_offsets=(bucket_offsets[b0:b1 + 1] - p0).astype(np.uint32),
                bucket_times=np.asarray(bucket_times, np.float64)[b0:b1], lo=lo, hi=hi)
