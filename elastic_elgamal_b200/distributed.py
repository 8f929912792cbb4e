"""Sharding of a ballot batch across the GPUs of one node (DESIGN.md "Multi-GPU").

Ballots are independent (examples/voting.rs:189-204 is a plain loop with a commutative fold), so rank g of P verifies
the contiguous slice [g*n/P, (g+1)*n/P) with its own Engine; verdict slices need no exchange.  The only exchange step
is the per-rank partial tally (options x 64 B): one all_gather followed by a local point addition on the GPU
(`Engine.ciphertexts_sum`) -- a collective reduce cannot be used because NCCL has no elliptic-curve operator.

The module only needs `torch.distributed` to be initialised by the caller (NCCL on the GPU box, gloo in the CPU
tests); it never creates a process group itself.
"""
import numpy as np


def shard_bounds(n, world, rank):
    """Contiguous slice of `n` items owned by `rank` out of `world` (sizes differ by at most one)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("rank/world out of range")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_partial_tallies(partial, dist=None, device=None):
    """all_gather of this rank's partial tally (options x 64 bytes) -> array (world, options, 64)."""
    partial = np.ascontiguousarray(partial, dtype=np.uint8)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return partial[None]
    import torch
    world = dist.get_world_size()
    t = torch.from_numpy(partial.reshape(-1).copy())      # flat: gloo and nccl both accept the 1-D form
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t)
    return out.cpu().numpy().reshape((world,) + partial.shape)


def gather_verdicts(verdicts, n, dist=None, device=None):
    """Concatenates the per-rank verdict slices (sizes from `shard_bounds`) in rank order on every rank."""
    verdicts = np.ascontiguousarray(verdicts, dtype=np.uint8)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return verdicts
    import torch
    world = dist.get_world_size()
    sizes = [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]
    width = max(sizes) if sizes else 0
    pad = np.zeros(width, np.uint8)
    pad[:verdicts.shape[0]] = verdicts
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * width, dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t)
    out = out.cpu().numpy().reshape(world, width)
    return np.concatenate([out[r, :sizes[r]] for r in range(world)]) if world else verdicts


def verify_choice_sharded(engine, options, choices, rings, sums=None, single=True, dist=None, device=None):
    """Drop-in for `for b in ballots: b.verify(&params); totals += b` over a batch that every rank holds (or can
    slice): each rank verifies its slice, then partial tallies are gathered and added.  Returns
    (local verdicts, (lo, hi), total tally) -- the tally is identical on every rank."""
    choices = np.asarray(choices, dtype=np.uint8).reshape(-1, options, 64)
    n = choices.shape[0]
    rings = np.asarray(rings, dtype=np.uint8).reshape(n, 1 + 2 * options, 32)
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    lo, hi = shard_bounds(n, world, rank)
    s = None if sums is None else np.asarray(sums, dtype=np.uint8).reshape(n, 64)[lo:hi]
    v, partial = engine.verify_choice(options, choices[lo:hi], rings[lo:hi], s, single=single, tally=True)
    parts = gather_partial_tallies(partial, dist, device)
    if parts.shape[0] == 1:
        return v, (lo, hi), parts[0]
    total, ok = engine.ciphertexts_sum(parts)
    if not ok:
        raise RuntimeError("a gathered partial tally does not decode (corrupted exchange)")
    return v, (lo, hi), total
