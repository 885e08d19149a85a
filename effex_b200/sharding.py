"""Multi-GPU plumbing for the FX path: time sharding + the one collective.

The reference is single-GPU; this is the north_star's extension.
In the reference's semantics every `num_samp` block is independent (zero PFB
history, per-block DC mean, one output row), so a recording shards by
CONTIGUOUS BLOCK RANGES with no halo and no data-path collective: each rank
produces the rows of its own range.  The only exchange is for integrations
that span ranks (a whole-recording spectrum, the accumulated lag
cross-spectrum): one reduce of the small float64 accumulators
(N x {complex cross, auto0, auto1} + a frame count) to rank 0.

`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the plumbing: rendezvous, the exchange
of mailbox tokens, the 5-word all-reduce of byte sums and the halo all-gather of streaming mode.  The
reduce of the accumulators itself is the library's own (fx_comm_*: peer-memory stores over NVLink fused
into the integrate epilogue) once `attach_comm(engine)` has run; without it (CPU tests, or accumulators
that do not come from an engine) `reduce_accumulators` uses `dist.reduce`.
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


def attach_comm(engine, group=None, slot_bytes: int = 0) -> bool:
    """Create the engine's mailbox, exchange the tokens over `group` and map the peers (every rank calls
    this).  Returns False -- and leaves the engine on the dist.reduce path -- when there is one rank only."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return False
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    tok = engine.comm_export(world, slot_bytes)
    mine = torch.frombuffer(bytearray(tok), dtype=torch.uint8).to(engine.tdev)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    ok = 1
    try:
        import os
        if os.environ.get("EFFEX_FX_NO_IPC"):          # test hook: behave like a box that refuses CUDA IPC
            raise RuntimeError("EFFEX_FX_NO_IPC is set")
        engine.comm_attach(rank, world, [bytes(g.cpu().numpy().tobytes()) for g in gathered])
    except Exception as e:             # e.g. CUDA IPC not permitted between these processes
        ok = 0
        engine.comm_error = str(e)
    # all ranks or none: a rank that could not map its peers takes everybody back to dist.reduce
    flag = torch.tensor([ok], dtype=torch.int32, device=engine.tdev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        engine.comm_world = 0
        return False
    return True


def reduce_accumulators(acc: dict, dst: int = 0, group=None, engine=None) -> dict:
    """One reduce(sum) of the per-integration accumulators to `dst` (in place
    on dst; other ranks' tensors are left unspecified by the collective)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return acc
    if engine is not None and engine.comm_attached and "flat" in acc:
        engine.reduce_inplace(acc["flat"], root=dst)       # fx_reduce_f64: peer-memory push + rank-ordered fold
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
    acc = reduce_accumulators(acc, dst=dst, group=group, engine=getattr(compute, "engine", None))
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
    compute.engine = engine
    return compute


def stream_integrate(engine, d_iq0: torch.Tensor, d_iq1: torch.Tensor, group=None, dst: int = 0):
    """Streaming-history integration of a recording that is time-sharded over the ranks of `group`
    (rank r holds the r-th contiguous slice, a whole number of blocks, in d_iq0/d_iq1).
    Exchanges: (1) all-reduce of the 4 byte sums -> recording-wide DC mean; (2) the PFB halo, i.e. the
    last (ntaps-1)*nbins samples of the left neighbour's slice (all-gather of the tiny tails);
    (3) one reduce of the float64 accumulators (the library's own when attach_comm has run: the push is
    fused into the integrate epilogue).  Returns the reduced accumulators (valid on dst).

    Preconditions (checked): the tensors hold exactly a whole number of blocks; every rank's slice is a
    whole number of FRAMES, so that the frame grid of the recording is the union of the ranks' grids
    (a ragged slice would shift every later rank's frames and drop its own tail -- not the
    one-giant-block semantics); a slice is at least as long as the halo."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    S, N, T = engine.num_samp, engine.nbins, engine.ntaps
    if d_iq0.numel() != d_iq1.numel() or d_iq0.numel() % (2 * S) != 0 or d_iq0.numel() == 0:
        raise ValueError("each channel must hold the same whole number (>= 1) of blocks of 2*num_samp bytes")
    n_blocks = d_iq0.numel() // (2 * S)
    if n_blocks > getattr(engine, "max_blocks", n_blocks):
        raise ValueError(f"slice of {n_blocks} blocks exceeds the engine's max_blocks = {engine.max_blocks}")
    span = n_blocks * S
    hb = 2 * (T - 1) * N
    if world > 1 and span % N != 0:
        raise ValueError(f"a rank's slice ({span} samples) must be a whole number of frames of {N} samples")
    if world > 1 and 2 * span < hb:
        raise ValueError("a rank's slice is shorter than the PFB halo")
    sums = torch.from_numpy(engine.span_sums(d_iq0, d_iq1, n_blocks).astype(np.int64)).to(d_iq0.device)
    count = torch.tensor([span], dtype=torch.int64, device=d_iq0.device)
    halo0 = halo1 = None
    if world > 1:
        both = torch.cat([sums, count])
        dist.all_reduce(both, op=dist.ReduceOp.SUM, group=group)
        sums, count = both[:4], both[4:]
        end = 2 * span                       # the halo is the tail of the SPAN
        tails = torch.stack([d_iq0[end - hb:end], d_iq1[end - hb:end]]).contiguous()
        gathered = [torch.empty_like(tails) for _ in range(world)]
        dist.all_gather(gathered, tails, group=group)
        if rank > 0:
            halo0, halo1 = gathered[rank - 1][0].contiguous(), gathered[rank - 1][1].contiguous()
    acc = engine.new_accumulators()
    h_sums, total = sums.cpu().numpy().astype(np.uint64), int(count.item())
    if world > 1 and getattr(engine, "comm_attached", False):
        engine.integrate_stream_reduce(d_iq0, d_iq1, acc, n_blocks, halo0, halo1, h_sums, total, root=dst)
        engine.sync()
        return acc
    engine.integrate_stream(d_iq0, d_iq1, acc, n_blocks, halo0, halo1, h_sums, total)
    return reduce_accumulators(acc, dst=dst, group=group)


def sharded_lag(engine, d_iq0: torch.Tensor, d_iq1: torch.Tensor, group=None, dst: int = 0):
    """Delay calibration over a time-sharded recording (reference effex.py:583-627 on one
    block): every rank accumulates FFT(a)*conj(FFT(b)) over ITS blocks, ONE reduce of the 2n-point
    cross-spectrum (complex64[M], 4 MiB at n = 2^18) to `dst`, then the inverse transform and the argmax
    on `dst`.  `engine` is a lag engine (num_samp = n; see Correlator._estimate_delay_gaussian).
    Returns (n, imax, xprev, xbest, xnext) on dst, None elsewhere."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_blocks = d_iq0.numel() // (2 * engine.num_samp)
    xacc = engine.lag_accumulate(d_iq0, d_iq1, n_blocks)
    if world > 1:
        if getattr(engine, "comm_attached", False):
            engine.reduce_inplace(xacc, root=dst)
        else:
            flat = torch.view_as_real(xacc)
            dist.reduce(flat, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if rank != dst:
        return None
    return engine.lag_finish(xacc)


def run_recording_sharded(cor, src0, src1, n_blocks: int | None = None, write_csv: bool = True, calibrate: bool = True,
                          group=None, dst: int = 0):
    """One spectrum/continuum run of `Correlator` over the ranks of `group` (one process per GPU): the
    drop-in's run loop (effex.py:326-417: first block pair calibrates and yields no row, every later block
    pair one row) with the blocks time-sharded by contiguous ranges.  `src0/src1` are two recordings every
    rank can read -- paths of regular files, or uint8 arrays / `synth.TiledRecording`s.  Rank `dst` calibrates
    on block 0 and broadcasts the delay, every rank turns its own block range into rows
    (`fx_process_host`), rank `dst` gathers them in block order and writes the reference's CSV.
    Returns the rows on `dst` (None elsewhere)."""
    import os
    from .ingest import RecordingReader, _is_regular_file
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if cor.mode == 'TEST':
        raise ValueError("TEST mode sweeps the delay block by block; it does not shard")
    S, N = int(cor.num_samp), int(cor.nbins)
    files = _is_regular_file(src0) and _is_regular_file(src1)
    have = (min(os.path.getsize(src0), os.path.getsize(src1)) if files else min(src0.size, src1.size)) // (2 * S)
    n_blocks = have if n_blocks is None else min(int(n_blocks), have)
    dev = torch.device("cuda", cor.device)

    def block(src, b0, nb):
        if files:
            return np.fromfile(src, dtype=np.uint8, count=2 * S * nb, offset=2 * S * b0)
        return src[2 * S * b0:2 * S * (b0 + nb)]
    first = 0
    if calibrate and n_blocks > 0:
        delay = torch.zeros(1, dtype=torch.float64, device=dev)
        if rank == dst:
            cor.gpu_iq_0 = torch.from_numpy(np.ascontiguousarray(block(src0, 0, 1))).to(dev)
            cor.gpu_iq_1 = torch.from_numpy(np.ascontiguousarray(block(src1, 0, 1))).to(dev)
            cor._calibrate_task()
            delay[0] = cor.calibrated_delay
        if world > 1:
            dist.broadcast(delay, src=dst, group=group)
        cor.calibrated_delay = float(delay.item())
        first = 1
    n_rows = max(n_blocks - first, 0)
    start, count = shard_range(n_rows, world, rank)
    eng = cor._main_engine(max_blocks=max(1, min(cor.batch_blocks, max(count, 1))))
    rows = np.empty((count, N), dtype=np.complex64)
    if files and count:
        reader = RecordingReader(src0, src1, S, batch_blocks=cor.batch_blocks, skip_blocks=first + start, max_blocks=count)
        for raw0, raw1, b0, nb in reader:
            eng.process_host(raw0, raw1, nb, out=rows[b0:b0 + nb])
    else:
        for b0 in range(0, count, cor.batch_blocks):
            nb = min(cor.batch_blocks, count - b0)
            eng.process_host(np.ascontiguousarray(block(src0, first + start + b0, nb)),
                             np.ascontiguousarray(block(src1, first + start + b0, nb)), nb, out=rows[b0:b0 + nb])
    drows = torch.from_numpy(rows).to(dev)
    allrows = gather_rows(drows, n_rows, dst=dst, group=group) if world > 1 else drows
    if rank != dst:
        return None
    out = allrows.cpu().numpy()
    if cor.mode == 'CONTINUUM':
        out = out.astype(np.complex128).mean(axis=1) / cor.bandwidth
    if write_csv:
        cor._write_metadata()
        writer = cor._start_writer()
        try:
            data = out.reshape(-1, 1) if cor.mode == 'CONTINUUM' else out
            for b0 in range(0, len(data), 256):
                writer.put(data[b0:b0 + 256])
        finally:
            writer.close()
    return out
