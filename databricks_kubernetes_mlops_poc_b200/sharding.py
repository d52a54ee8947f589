"""Row sharding across the GPUs of one box (host-side bookkeeping).

Rows are independent and the forest is replicated, so the predict path needs no collective: a batch is
cut into contiguous slices, slice i goes to GPU i (``b2f_predict_multi`` in the C ABI does exactly
``[n*i/k, n*(i+1)/k)``), results land in disjoint slices of one output buffer.  The reference's analogue
is the k8s Service spreading requests over pod replicas (reference ``kubernetes/manifest.yml:23-36``).
The only cross-GPU exchange is the drift-monitor moments merge: every rank contributes one
(count, mean, M2) triple per feature (576 bytes) and merges them with Chan's formula.
"""

from __future__ import annotations

import numpy as np

from .engine import moments_merge


def shard_bounds(n: int, k: int):
    """Contiguous, balanced slices: [(lo, hi)] * k with sizes differing by at most one row."""
    return [(n * i // k, n * (i + 1) // k) for i in range(k)]


def round_robin_batches(n_rows: int, batch: int, k: int):
    """Stream of whole batches dealt round-robin to k GPUs: yields (gpu, lo, hi)."""
    for b, lo in enumerate(range(0, n_rows, batch)):
        yield b % k, lo, min(lo + batch, n_rows)


def allgather_merge_moments(local: np.ndarray, dist=None) -> np.ndarray:
    """Merge per-rank (24, 3) moment triples across a torch.distributed group (gloo on CPU boxes,
    nccl on GPU boxes).  ``b2f_moments_allgather`` is the same thing through the engine's own NCCL
    communicator."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local, dtype=np.float64)
    import torch

    t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64).reshape(-1))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    return moments_merge(np.stack([p.cpu().numpy() for p in parts]))
