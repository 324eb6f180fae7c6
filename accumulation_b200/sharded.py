"""Point-range sharding of one MSM across the GPUs of a box (north_star): one process per GPU,
each owns the contiguous slice [g*n/G, (g+1)*n/G) of the commitment key (resident in its HBM) and of the
scalars; the only exchange is one all-gather of a 128-byte XYZZ partial per GPU per MSM over NCCL/NVLink,
after which rank 0 adds the G partials and normalises.  torch.distributed is the plumbing only.

The h(X) coefficients of the IPA decider need no exchange at all: every GPU expands its own index range from
the k challenges."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

PARTIAL_WORDS = 16   # X, Y, ZZ, ZZZ: 4 x 4 u64


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """(start, count) of rank's contiguous point range; ranges tile [0, n) exactly, sizes differ by <= 1."""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi - lo


def gather_partials(local_partial, world: int, group=None):
    """all-gather of one 16-word partial per rank -> (world, 16) tensor in rank order, on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_partial.reshape(1, PARTIAL_WORDS)
    out = torch.empty((world, PARTIAL_WORDS), dtype=local_partial.dtype, device=local_partial.device)
    dist.all_gather_into_tensor(out, local_partial.reshape(1, PARTIAL_WORDS), group=group)
    return out


class ShardedMSM:
    """This rank's share of a sharded commitment key plus the gather/combine step.

    `bases_slice_xy` are the bases of THIS rank's range (shard_range(n_total, rank, world))."""

    def __init__(self, ctx, curve: int, bases_slice, n_total: int, rank: int = 0, world: int = 1, group=None,
                 device: Optional[str] = None):
        """bases_slice: (count, 8) uint64 array of this rank's bases, or an already registered `Bases` handle
        (e.g. from Context.register_synthetic_bases(first_index=start))."""
        import torch
        self.ctx, self.curve, self.rank, self.world, self.group = ctx, curve, rank, world, group
        self.n_total = n_total
        self.start, self.count = shard_range(n_total, rank, world)
        if hasattr(bases_slice, "handle"):
            self.bases = bases_slice
            have = bases_slice.n
        else:
            xy = np.ascontiguousarray(bases_slice, dtype=np.uint64).reshape(-1, 8)
            have = xy.shape[0]
            self.bases = ctx.register_bases(curve, xy) if have == self.count else None   # a handle may hold spare bases
        if have < self.count or (have != self.count and self.bases is None):
            raise ValueError(f"rank {rank} owns {self.count} bases, got {have}")
        self.device = device or f"cuda:{torch.cuda.current_device()}"
        self.partial = torch.zeros(PARTIAL_WORDS, dtype=torch.int64, device=self.device)

    def _finish(self, stream_ptr: int):
        """gather the partials and combine on rank 0 -> (xy, inf) on rank 0, None elsewhere"""
        allp = gather_partials(self.partial, self.world, self.group)
        if self.rank != 0:
            return None
        return self.ctx.combine_partials_dev(self.curve, allp.data_ptr(), self.world, stream=stream_ptr)

    def msm_dev(self, d_scalars, n: Optional[int] = None, montgomery: bool = True):
        """d_scalars: torch int64 CUDA tensor holding this rank's (count, 4) scalar slice."""
        import torch
        st = torch.cuda.current_stream().cuda_stream
        n = self.count if n is None else n
        if self.world == 1:      # nothing to gather: normalise inside the same call
            return self.ctx.msm_dev(self.bases, d_scalars.data_ptr(), n, montgomery=montgomery, stream=st)
        self.ctx.msm_partial_dev(self.bases, d_scalars.data_ptr(), n, self.partial.data_ptr(), montgomery=montgomery, stream=st)
        return self._finish(st)

    def msm_host(self, h_scalars, d_staging, montgomery: bool = True):
        """h_scalars: pinned torch int64 tensor with this rank's scalar slice; d_staging: CUDA tensor of the same
        shape.  H2D, local MSM, gather, combine; returns (xy, inf) on rank 0."""
        d_staging.copy_(h_scalars, non_blocking=True)
        return self.msm_dev(d_staging, montgomery=montgomery)

    def ipa_final_key(self, challenges_mont, k: int):
        """final_key = cm_commit(key, h.compute_coeffs()) with the key sharded by point range: each GPU expands
        coefficients [start, start + count) of h(X) on the fly; no data-path collective besides the gather."""
        import torch
        st = torch.cuda.current_stream().cuda_stream
        if (1 << k) != self.n_total:
            raise ValueError("sharded ipa_final_key expects a key of exactly 2^k bases")
        self.ctx.ipa_final_key_partial_dev(self.bases, challenges_mont, self.start, self.count, self.partial.data_ptr(), stream=st)
        return self._finish(st)

    def release(self):
        self.bases.release()


def sharded_sum_check(n: int, world: int, partial_fn: Callable[[int, int], object]):
    """Host-logic helper used by the gloo CPU tests: evaluates partial_fn(start, count) for every rank's range."""
    return [partial_fn(*shard_range(n, r, world)) for r in range(world)]
