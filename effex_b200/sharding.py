"""Multi-GPU plumbing for the FX path: time sharding + the one collective.

The reference is single-GPU; this is the north_star's extension.
In the reference's semantics every `num_samp` block is independent (zero PFB
history, per-block DC mean, one output row), so a recording shards by
CONTIGUOUS BLOCK RANGES with no halo and no data-path collective: each rank
produces the rows of its own range.  The only exchange is for integrations
that span ranks (a whole-recording spectrum, the accumulated lag
cross-spectrum): one reduce of the small float64 accumulators
(N x {complex cross, auto0, auto1} + a frame count) to rank 0.

`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the transport.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

ACC_KEYS = ("x", "a0", "a1", "frames")


def shard_range(n_blocks: int, world: int, rank: int):
    """Contiguous block range [start, start+count) of `rank`; the first
    n_blocks % world ranks take one extra block."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(int(n_blocks), world)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def reduce_accumulators(acc: dict, dst: int = 0, group=None) -> dict:
    """One reduce(sum) of the per-integration accumulators to `dst` (in place
    on dst; other ranks' tensors are left unspecified by the collective)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return acc
    if "flat" in acc:                       # FxEngine.new_accumulators: already one buffer
        dist.reduce(acc["flat"], dst=dst, op=dist.ReduceOp.SUM, group=group)
        return acc
    flat = torch.cat([acc[k].reshape(-1) for k in ACC_KEYS])
    dist.reduce(flat, dst=dst, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for k in ACC_KEYS:
        n = acc[k].numel()
        acc[k].copy_(flat[off:off + n].reshape(acc[k].shape))
        off += n
    return acc


def gather_rows(rows: torch.Tensor, n_blocks: int, dst: int = 0, group=None):
    """Concatenate every rank's rows in block order on `dst` (None elsewhere).
    rows: [count_of_this_rank, N] complex64."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return rows
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nbins = rows.shape[1]
    maxc = shard_range(n_blocks, world, 0)[1]
    pad = torch.zeros((maxc, nbins, 2), dtype=torch.float32, device=rows.device)
    pad[: rows.shape[0]] = torch.view_as_real(rows)
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    parts = [torch.view_as_complex(bufs[r][: shard_range(n_blocks, world, r)[1]].contiguous()) for r in range(world)]
    return torch.cat(parts, dim=0)


def finish_integration(acc: dict, rot=None):
    """1/frames, conj(rot), fftshift in float64 on the host (effex.py:519-521)."""
    frames = float(acc["frames"].reshape(-1)[0].item())
    x = acc["x"].detach().cpu().numpy().astype(np.float64).view(np.complex128) / frames
    if rot is not None:
        x = x * np.conj(rot)
    a0 = acc["a0"].detach().cpu().numpy() / frames
    a1 = acc["a1"].detach().cpu().numpy() / frames
    return np.fft.fftshift(x), np.fft.fftshift(a0), np.fft.fftshift(a1)


def sharded_run(compute, n_blocks: int, group=None, dst: int = 0):
    """Drive one sharded pass.  `compute(start, count)` runs this rank's block
    range and returns (rows [count, N] complex64, acc dict of float64 tensors).
    Returns (all rows on dst or None, reduced acc)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    start, count = shard_range(n_blocks, world, rank)
    rows, acc = compute(start, count)
    acc = reduce_accumulators(acc, dst=dst, group=group)
    rows = gather_rows(rows, n_blocks, dst=dst, group=group)
    return rows, acc


def engine_compute(engine, d_iq0: torch.Tensor, d_iq1: torch.Tensor):
    """compute() for sharded_run on top of an FxEngine and the full recording
    resident on this rank's device (each rank touches only its own range)."""
    S = engine.num_samp

    def compute(start, count):
        acc = engine.new_accumulators()
        lo, hi = 2 * S * start, 2 * S * (start + count)
        rows = engine.process(d_iq0[lo:hi], d_iq1[lo:hi], count, acc=acc)
        return rows, acc
    return compute


def stream_integrate(engine, d_iq0: torch.Tensor, d_iq1: torch.Tensor, group=None, dst: int = 0):
    """Streaming-history integration of a recording that is time-sharded over the ranks of `group`
    (rank r holds the r-th contiguous slice, a whole number of blocks, in d_iq0/d_iq1).
    Exchanges: (1) all-reduce of the 4 byte sums -> recording-wide DC mean; (2) the PFB halo, i.e. the
    last (ntaps-1)*nbins samples of the left neighbour's slice (all-gather of the tiny tails);
    (3) one reduce of the float64 accumulators.  Returns the reduced accumulators (valid on dst)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_blocks = d_iq0.numel() // (2 * engine.num_samp)
    sums = torch.from_numpy(engine.span_sums(d_iq0, d_iq1, n_blocks).astype(np.int64)).to(d_iq0.device)
    count = torch.tensor([n_blocks * engine.num_samp], dtype=torch.int64, device=d_iq0.device)
    halo0 = halo1 = None
    if world > 1:
        both = torch.cat([sums, count])
        dist.all_reduce(both, op=dist.ReduceOp.SUM, group=group)
        sums, count = both[:4], both[4:]
        hb = 2 * (engine.ntaps - 1) * engine.nbins
        tails = torch.stack([d_iq0[-hb:], d_iq1[-hb:]]).contiguous()
        gathered = [torch.empty_like(tails) for _ in range(world)]
        dist.all_gather(gathered, tails, group=group)
        if rank > 0:
            halo0, halo1 = gathered[rank - 1][0].contiguous(), gathered[rank - 1][1].contiguous()
    acc = engine.new_accumulators()
    engine.integrate_stream(d_iq0, d_iq1, acc, n_blocks, halo0, halo1, sums.cpu().numpy().astype(np.uint64),
                            int(count.item()))
    return reduce_accumulators(acc, dst=dst, group=group)
