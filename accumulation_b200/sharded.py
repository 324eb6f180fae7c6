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


def _stream_handle() -> int:
    """cudaStream_t of torch's current stream for the *_dev entry points.  Those read buffers torch (NCCL) has just
    produced, so the work must be ordered on torch's stream -- but torch's default stream has the handle 0, which the
    C-ABI defines as "the ctx stream" (a non-blocking stream that does NOT wait for the default stream).  The legacy
    default stream is therefore passed by its explicit handle cudaStreamLegacy = 0x1."""
    import torch
    return torch.cuda.current_stream().cuda_stream or 1


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
        st = _stream_handle()
        n = self.count if n is None else n
        if self.world == 1:      # nothing to gather: normalise inside the same call
            return self.ctx.msm_dev(self.bases, d_scalars.data_ptr(), n, montgomery=montgomery, stream=st)
        self.ctx.msm_partial_dev(self.bases, d_scalars.data_ptr(), n, self.partial.data_ptr(), montgomery=montgomery, stream=st)
        return self._finish(st)

    def msm_host(self, h_scalars, d_staging=None, montgomery: bool = True):
        """h_scalars: pinned torch int64 tensor (or numpy array) with this rank's scalar slice in HOST memory.  The same
        reference-facing C-ABI path as the single-GPU call: the library uploads the slice (second point segment in flight
        while the first is accumulated) and leaves this rank's partial on the device (accmsm_msm_partial); then gather +
        combine.  Returns (xy, inf) on rank 0.  d_staging is unused (kept for callers of the first version)."""
        if self.world == 1:
            arr = h_scalars.numpy().view(np.uint64) if hasattr(h_scalars, "numpy") else h_scalars
            return self.ctx.msm(self.bases, arr, montgomery=montgomery, n=self.count)
        ptr = h_scalars.data_ptr() if hasattr(h_scalars, "data_ptr") else np.ascontiguousarray(h_scalars, dtype=np.uint64).ctypes.data
        self.ctx.msm_partial(self.bases, int(ptr), self.partial.data_ptr(), montgomery=montgomery, n=self.count)
        return self._finish(_stream_handle())

    def ipa_final_key(self, challenges_mont, k: int):
        """final_key = cm_commit(key, h.compute_coeffs()) with the key sharded by point range: each GPU expands
        coefficients [start, start + count) of h(X) on the fly; no data-path collective besides the gather."""
        import torch
        st = _stream_handle()
        if (1 << k) != self.n_total:
            raise ValueError("sharded ipa_final_key expects a key of exactly 2^k bases")
        self.ctx.ipa_final_key_partial_dev(self.bases, challenges_mont, self.start, self.count, self.partial.data_ptr(), stream=st)
        return self._finish(st)

    def release(self):
        self.bases.release()


def sharded_sum_check(n: int, world: int, partial_fn: Callable[[int, int], object]):
    """Host-logic helper used by the gloo CPU tests: evaluates partial_fn(start, count) for every rank's range."""
    return [partial_fn(*shard_range(n, r, world)) for r in range(world)]


# ---------------------------------------------------------------------------------------------------------------
# IpaPC::open across GPUs ("IPA open folding"): cyclic sharding
# ---------------------------------------------------------------------------------------------------------------

def cyclic_shard(a, rank: int, world: int):
    """rows rank, rank + world, rank + 2 world, ... : the fold partners i and i + n/2 of every round share a rank
    while n/2 >= world"""
    return np.ascontiguousarray(np.asarray(a)[rank::world])


def gather_rows(local_row, world: int, group=None):
    """all-gather of one equally sized int64 row per rank -> (world, len) tensor in rank order, on every rank"""
    import torch
    import torch.distributed as dist
    row = local_row.reshape(1, -1)
    if world == 1:
        return row
    out = torch.empty((world, row.shape[1]), dtype=row.dtype, device=row.device)
    dist.all_gather_into_tensor(out, row.contiguous(), group=group)
    return out


def z_fold_factor(field: int, point_mont, challenges_mont, n0: int) -> np.ndarray:
    """After r rounds the folded z-vector of an opening at `point` is F_r * (1, z, z^2, ...) with
    F_r = prod_{j=1..r} (1 + xi_j z^(n0 / 2^j))  (z_l += xi z_r, round by round).  Montgomery limbs in and out."""
    from .mirror import _MODULI, _fe_to_int, _int_to_fe
    m = _MODULI[field]
    z = _fe_to_int(field, point_mont)
    f, h = 1, n0
    for xi in np.asarray(challenges_mont, dtype=np.uint64).reshape(-1, 4):
        h //= 2
        f = f * (1 + _fe_to_int(field, xi) * pow(z, h, m)) % m
    return _int_to_fe(field, f)


class ShardedIpaOpen:
    """One rank's share of IpaPC::open over a commitment key sharded cyclically across `world` = 2^g GPUs.

    Rank r registers key[r::world] (+ the hiding generator as the base after them when h' = xi_0 * h is given by
    xi_0) and holds coeffs[r::world].  The first k - g rounds run as an ordinary opening session of length 2^k / world
    on every GPU; per round the only exchange is one all-gather of this rank's 2 x 128-byte shares of (l, r).  The last
    g rounds are an opening of length `world` over the gathered (final key, coefficient) pairs and run replicated.
    The result is bit-identical to the single-GPU session (and to the oracle's round-by-round folding)."""

    def __init__(self, ctx, curve: int, key_shard, k: int, rank: int = 0, world: int = 1, group=None,
                 hiding_index: Optional[int] = None, device: Optional[str] = None):
        import torch
        if world < 1 or world & (world - 1):
            raise ValueError("world must be a power of two")
        self.log_world = world.bit_length() - 1
        if k < self.log_world:
            raise ValueError("the opening must have at least one coefficient per rank")
        self.ctx, self.curve, self.key, self.k, self.rank, self.world, self.group = ctx, curve, key_shard, k, rank, world, group
        self.k_local = k - self.log_world
        self.hiding_index = hiding_index
        self.device = device or f"cuda:{torch.cuda.current_device()}"
        self.partials = torch.zeros(2 * PARTIAL_WORDS, dtype=torch.int64, device=self.device)
        self.session = None

    # -- the steps of one rank (tests drive several virtual ranks on one GPU through these)
    def begin(self, coeffs_shard_mont, point_mont, h_prime_xy=None, xi0_mont=None):
        self.session = self.ctx.ipa_open_begin_shard(self.key, coeffs_shard_mont, self.k_local, point_mont,
                                                     None if xi0_mont is not None else h_prime_xy,
                                                     shard_index=self.rank, log_shards=self.log_world)
        if xi0_mont is not None:
            if self.hiding_index is None:
                raise ValueError("xi0 needs hiding_index (the hiding generator inside this rank's key)")
            self.ctx.ipa_open_use_hiding_generator(self.session, self.hiding_index, xi0_mont)

    def round_partials(self):
        """this rank's un-normalised shares of (l, r) as a (32,) int64 CUDA tensor"""
        self.ctx.ipa_open_round_partial_dev(self.session, self.partials.data_ptr())
        return self.partials

    def combine(self, all_partials):
        """all_partials: (world, 32) gathered shares -> ((l_xy, l_inf), (r_xy, r_inf)); every rank gets the same"""
        ap = all_partials.contiguous()
        out = self.ctx.combine_partials_batch_dev(self.curve, ap.data_ptr(), self.world, 2, stream=_stream_handle())
        return out[0], out[1]

    def fold(self, xi_mont, xi_inv_mont):
        self.ctx.ipa_open_fold(self.session, xi_mont, xi_inv_mont)

    def finish_local(self):
        """(final key of this rank's shard, its last coefficient): element `rank` of the length-`world` opening left"""
        fk, c = self.ctx.ipa_open_finish(self.session)
        self.session = None
        return fk, c

    def tail(self, fk_all, c_all, point_mont, challenges_mont, round_challenge, prev_xi, h_prime_xy=None, xi0_mont=None,
             hiding_generator_xy=None):
        """the last log2(world) rounds on the gathered pairs (replicated on every rank; needs no exchange)"""
        from . import scalar_field
        from .mirror import _MODULI, _fe_to_int, _int_to_fe
        field = scalar_field(self.curve)
        key_xy = np.ascontiguousarray(fk_all, dtype=np.uint64).reshape(self.world, 8)
        if xi0_mont is not None:
            key_xy = np.concatenate([key_xy, np.asarray(hiding_generator_xy, dtype=np.uint64).reshape(1, 8)])
        tiny = self.ctx.register_bases(self.curve, key_xy)
        try:
            scale = z_fold_factor(field, point_mont, challenges_mont, 1 << self.k)
            sess = self.ctx.ipa_open_begin_shard(tiny, c_all, self.log_world, point_mont,
                                                 None if xi0_mont is not None else h_prime_xy, z_scale_mont=scale)
            if xi0_mont is not None:
                self.ctx.ipa_open_use_hiding_generator(sess, self.world, xi0_mont)
            l_vec, r_vec, chs, xi = [], [], [], prev_xi
            for _ in range(self.log_world):
                l, r = self.ctx.ipa_open_round(sess)
                xi = np.ascontiguousarray(round_challenge(xi, l, r), dtype=np.uint64).reshape(4)
                self.ctx.ipa_open_fold(sess, xi, _int_to_fe(field, pow(_fe_to_int(field, xi), -1, _MODULI[field])))
                l_vec.append(l); r_vec.append(r); chs.append(xi)
            fk, c = self.ctx.ipa_open_finish(sess)
        finally:
            tiny.release()
        return l_vec, r_vec, fk, c, chs

    # -- the whole opening of this rank, exchanges over torch.distributed (NCCL)
    def open(self, coeffs_shard_mont, point_mont, round_challenge, h_prime_xy=None, xi0_mont=None):
        """-> (l_vec, r_vec, final_comm_key_xy, c, challenges), identical on every rank.  `round_challenge(prev, l, r)`
        is the host sponge; it must be deterministic (every rank evaluates it on the same (l, r))."""
        import torch
        from . import scalar_field
        from .mirror import _MODULI, _fe_to_int, _int_to_fe
        field = scalar_field(self.curve)
        self.begin(coeffs_shard_mont, point_mont, h_prime_xy, xi0_mont)
        l_vec, r_vec, chs, xi = [], [], [], None
        for _ in range(self.k_local):
            allp = gather_rows(self.round_partials(), self.world, self.group)
            l, r = self.combine(allp)
            xi = np.ascontiguousarray(round_challenge(xi, l, r), dtype=np.uint64).reshape(4)
            self.fold(xi, _int_to_fe(field, pow(_fe_to_int(field, xi), -1, _MODULI[field])))
            l_vec.append(l); r_vec.append(r); chs.append(xi)
        fk, c = self.finish_local()
        if self.world == 1:
            return l_vec, r_vec, fk, c, chs
        row = torch.from_numpy(np.concatenate([fk, c]).view(np.int64).copy()).to(self.device)
        pairs = gather_rows(row, self.world, self.group).cpu().numpy().view(np.uint64)
        hg = self.ctx.download_bases(self.key, self.hiding_index, 1) if xi0_mont is not None else None
        tl, tr, fk, c, tch = self.tail(pairs[:, :8], pairs[:, 8:], point_mont, chs, round_challenge, xi, h_prime_xy, xi0_mont, hg)
        return l_vec + tl, r_vec + tr, fk, c, chs + tch
