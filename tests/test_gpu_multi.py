"""The N>1 path on hardware: contiguous time shards, PFB halos + recording-wide byte sums in
streaming mode, the reduce of the float64 accumulators through the library's own mailboxes (fx_comm_*),
and the multi-GPU lag search -- one process per GPU under NCCL at world sizes 2 and 4 (skipped when the box
has fewer GPUs), against the float64 oracle on the whole recording.

The mailbox protocol itself (epochs, parity slots, per-CTA flags, back-pressure) also runs on ONE GPU:
two handles on the same device in one process are two ranks of a world of 2.
"""
import os

import numpy as np
import pytest
import torch

from oracle import fx_oracle as orc
from effex_b200 import synth, sharding
from effex_b200.engine import FxEngine

pytestmark = pytest.mark.gpu
TOL = 1e-4
BW, FC = 2.4e6, 1.4204e9


def dev(a, d=0):
    return torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{d}")


def close(got, ref, tol=TOL):
    got = np.asarray(got, dtype=np.complex128); ref = np.asarray(ref, dtype=np.complex128)
    return (np.abs(got - ref).max() <= tol * np.abs(ref).max()
            and np.linalg.norm(got - ref) <= tol * np.linalg.norm(ref))


def _attach_local(engines):
    """ranks of one process: exchange tokens by hand"""
    world = len(engines)
    toks = [e.comm_export(world) for e in engines]
    for r, e in enumerate(engines):
        e.comm_attach(r, world, toks)


# ---- one GPU: the protocol ------------------------------------------------------------------------------
def test_reduce_world_of_one_equals_integrate():
    S, N, nb = 2**15, 4096, 6
    raw0, raw1 = synth.correlated_pair(nb * S, delay=5, seed=3)
    d0, d1 = dev(raw0), dev(raw1)
    eng = FxEngine(S, N, 4, max_blocks=nb)
    _attach_local([eng])
    ref = eng.new_accumulators()
    rows_ref = eng.process(d0, d1, nb, acc=ref)
    acc = eng.new_accumulators()
    rows = eng.process_reduce(d0, d1, nb, acc=acc, root=0)
    eng.sync()
    assert torch.equal(rows, rows_ref)
    assert torch.equal(acc["flat"], ref["flat"])           # same sums, same order: bit-identical
    eng.close()


def test_cross_only_accumulators_skip_the_auto_powers():
    """FX_FLAG_CROSS_ONLY: the cross-spectrum and frame-count parts are bit-identical to the full
    accumulators, the auto-power parts stay untouched; rows are the same rows."""
    S, N, nb = 2**15, 4096, 5
    raw0, raw1 = synth.correlated_pair(nb * S, delay=5, seed=3)
    d0, d1 = dev(raw0), dev(raw1)
    full, lean = FxEngine(S, N, 4, max_blocks=nb), FxEngine(S, N, 4, max_blocks=nb, cross_only=True)
    a, b = full.new_accumulators(), lean.new_accumulators()
    b["a0"].fill_(7.0)
    ra, rb = full.process(d0, d1, nb, acc=a), lean.process(d0, d1, nb, acc=b)
    full.integrate(d0, d1, a, nb); lean.integrate(d0, d1, b, nb)
    full.sync(); lean.sync()
    assert torch.equal(ra, rb)
    assert torch.equal(a["x"], b["x"]) and a["frames"].item() == b["frames"].item() == 2 * nb * (S // N)
    assert torch.all(b["a0"] == 7.0) and torch.all(b["a1"] == 0.0) and torch.all(a["a0"] > 0)
    full.close(); lean.close()


@pytest.mark.parametrize("N,S", [(4096, 2**15), (1024, 2**13), (8192, 2**16), (128, 2**12)])
def test_two_ranks_on_one_gpu_pipelined_epochs(N, S):
    """Two handles = two ranks.  Six reduce epochs are issued back to back without a host sync (parity
    slots are reused from epoch 3 on: the push waits for the root's fold of epoch e-2), alternating which
    rank issues first; the root's accumulators must equal the sum of everything both ranks integrated."""
    nb = 4
    raws = [synth.correlated_pair(nb * S, delay=5, seed=30 + r) for r in range(2)]
    dv = [(dev(a), dev(b)) for a, b in raws]
    engs = [FxEngine(S, N, 4, max_blocks=nb) for _ in range(2)]
    _attach_local(engs)
    acc = engs[0].new_accumulators()
    epochs = 6
    for e in range(epochs):
        order = (0, 1) if e % 2 == 0 else (1, 0)
        for r in order:
            engs[r].process_reduce(dv[r][0], dv[r][1], nb, acc=acc if r == 0 else None, root=0)
    for eng in engs:
        eng.sync()
    want = engs[0].new_accumulators()
    for r in range(2):
        part = engs[r].new_accumulators()
        engs[r].integrate(dv[r][0], dv[r][1], part, nb)
        engs[r].sync()
        want["flat"] += part["flat"]
    want["flat"] *= epochs
    assert acc["frames"].item() == 2 * epochs * nb * (S // N)
    np.testing.assert_allclose(acc["flat"].cpu().numpy(), want["flat"].cpu().numpy(), rtol=1e-12, atol=1e-9)
    # a non-zero root, and the generic in-place reduce (float32 / complex64)
    acc1 = engs[1].new_accumulators()
    for r in (0, 1):
        engs[r].process_reduce(dv[r][0], dv[r][1], nb, acc=acc1 if r == 1 else None, root=1)
    bufs = [torch.full((4 * N - 3,), float(r + 1), dtype=torch.float32, device="cuda") for r in range(2)]
    for r in (1, 0):
        engs[r].reduce_inplace(bufs[r], root=0)
    for eng in engs:
        eng.sync()
    np.testing.assert_allclose(acc1["flat"].cpu().numpy() * epochs, want["flat"].cpu().numpy(), rtol=1e-12, atol=1e-9)
    assert torch.all(bufs[0] == 3.0) and torch.all(bufs[1] == 2.0)
    for eng in engs:
        eng.close()


def test_missing_rank_times_out_instead_of_hanging(monkeypatch):
    from effex_b200.engine import FxCommError
    monkeypatch.setenv("EFFEX_FX_COMM_TIMEOUT_MS", "50")
    S, N, nb = 2**13, 1024, 2
    raw0, raw1 = synth.correlated_pair(nb * S, seed=1)
    engs = [FxEngine(S, N, 4, max_blocks=nb) for _ in range(2)]
    _attach_local(engs)
    acc = engs[0].new_accumulators()
    engs[0].process_reduce(dev(raw0), dev(raw1), nb, acc=acc, root=0)      # rank 1 never contributes
    with pytest.raises(FxCommError):
        engs[0].sync()
    for eng in engs:
        eng.close()


def test_sharded_lag_on_logical_shards():
    """accumulate per shard + sum + finish == the one-call lag search over all blocks (integer lag exact)."""
    S, nb = 2**16, 6
    raw0, raw1 = synth.correlated_pair(nb * S, delay=37, seed=12)
    d0, d1 = dev(raw0), dev(raw1)
    eng = FxEngine(S, 4096, 1, max_blocks=nb)
    n, imax, p, q, r = eng.lag(d0, d1, nb)
    assert n - imax == 37
    tot = None
    for rank in range(3):
        start, count = sharding.shard_range(nb, 3, rank)
        x = eng.lag_accumulate(d0[2 * S * start:2 * S * (start + count)], d1[2 * S * start:2 * S * (start + count)], count)
        tot = x if tot is None else tot + x
    n2, imax2, p2, q2, r2 = eng.lag_finish(tot)
    assert imax2 == imax
    assert (p2, q2, r2) == pytest.approx((p, q, r), rel=1e-5)
    eng.close()


# ---- one process per GPU under NCCL -----------------------------------------------------------------------
S_MP, N_MP, NB_MP = 2**15, 4096, 13           # 13 blocks: ragged over 2 and 4 ranks


def _recording():
    return synth.correlated_pair(NB_MP * S_MP, delay=37, dc0=0.02 + 0.01j, dc1=-0.015j, seed=77)


def _nccl_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    raw0, raw1 = _recording()
    tau = 37 / BW
    start, count = sharding.shard_range(NB_MP, world, rank)
    sl = slice(2 * S_MP * start, 2 * S_MP * (start + count))
    d0, d1 = dev(raw0[sl], rank), dev(raw1[sl], rank)
    eng = FxEngine(S_MP, N_MP, 4, device=rank, max_blocks=NB_MP)
    eng.set_delay(BW, FC, tau)
    assert sharding.attach_comm(eng)

    # (1) reference-exact mode: rows of the own range + reduce of the accumulators
    def compute(s, c):
        acc = eng.new_accumulators()
        return eng.process(d0, d1, c, acc=acc), acc
    compute.engine = eng
    rows, acc = sharding.sharded_run(compute, NB_MP)
    eng.sync()
    # (2) the fused variant, several epochs without a host sync: rows + push from the integrate epilogue
    acc2 = eng.new_accumulators()
    for _ in range(5):
        eng.process_reduce(d0, d1, count, acc=acc2, root=0)
    eng.sync()
    # (3) streaming mode: byte-sum all-reduce, halo all-gather, reduce fused into the integrate epilogue
    acc3 = sharding.stream_integrate(eng, d0, d1)
    eng.sync()
    # (4) lag search over the shards: reduce of the 2n-point cross-spectrum, inverse + argmax on rank 0
    leng = FxEngine(S_MP, 4096, 1, device=rank, max_blocks=NB_MP)
    assert sharding.attach_comm(leng, slot_bytes=8 * leng.lag_fft_len())
    lag = sharding.sharded_lag(leng, d0, d1)
    leng.sync()
    if rank == 0:
        x, a0, a1 = sharding.finish_integration(acc)
        x2, _, _ = sharding.finish_integration(acc2)
        x3, a03, _ = sharding.finish_integration(acc3)
        np.savez(out, rows=rows.cpu().numpy(), x=x, a0=a0, frames=acc["frames"].cpu().numpy(), x2=x2,
                 frames2=acc2["frames"].cpu().numpy(), x3=x3, a03=a03, frames3=acc3["frames"].cpu().numpy(),
                 lag=np.array(lag, dtype=np.float64))
    else:
        assert rows is None and lag is None
    dist.barrier()
    torch.cuda.synchronize()
    eng.close()
    leng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_nccl_time_shards_match_the_oracle(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "m.npz")
    port = 29600 + os.getpid() % 2000 + world
    mp.spawn(_nccl_worker, args=(world, port, out), nprocs=world, join=True)
    got = np.load(out)
    raw0, raw1 = _recording()
    tau = 37 / BW
    P = S_MP // N_MP
    ref_rows = orc.process_recording_u8(raw0, raw1, S_MP, N_MP, BW, FC, tau, 4, 0, NB_MP)
    for b in range(NB_MP):
        assert close(got["rows"][b], ref_rows[b]), b
    # accumulators: un-rotated mean over all frames of all blocks (reference mode: per-block mean, zero history)
    ref0 = orc.process_recording_u8(raw0, raw1, S_MP, N_MP, BW, FC, 0.0, 4, 0, NB_MP).mean(axis=0)
    assert got["frames"][0] == NB_MP * P and got["frames2"][0] == 5 * NB_MP * P
    assert close(got["x"], ref0) and close(got["x2"], ref0)
    # streaming mode == the reference's arithmetic on the recording as ONE block
    w = orc.pfb_window(4, N_MP)
    f0 = orc.spectrometer_poly(orc.block_from_u8(raw0), 4, N_MP, w)
    f1 = orc.spectrometer_poly(orc.block_from_u8(raw1), 4, N_MP, w)
    assert got["frames3"][0] == f0.shape[0]
    assert close(got["x3"], np.fft.fftshift((f0 * np.conj(f1)).mean(axis=0)))
    assert close(got["a03"], np.fft.fftshift((abs(f0) ** 2).mean(axis=0)))
    n, imax = int(got["lag"][0]), int(got["lag"][1])
    assert n - imax == 37


def test_two_gpus_one_process_peer_access():
    """same-process ranks on two devices: the mailbox is reached through cudaDeviceEnablePeerAccess"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    S, N, nb = 2**15, 4096, 4
    raws = [synth.correlated_pair(nb * S, delay=5, seed=40 + r) for r in range(2)]
    engs = [FxEngine(S, N, 4, device=r, max_blocks=nb) for r in range(2)]
    dv = [(dev(a, r), dev(b, r)) for r, (a, b) in enumerate(raws)]
    _attach_local(engs)
    acc = engs[0].new_accumulators()
    want = None
    for e in range(4):
        for r in ((0, 1) if e % 2 else (1, 0)):
            with torch.cuda.device(r):
                engs[r].process_reduce(dv[r][0], dv[r][1], nb, acc=acc if r == 0 else None, root=0)
    for r in range(2):
        with torch.cuda.device(r):
            engs[r].sync()
            part = engs[r].new_accumulators()
            engs[r].integrate(dv[r][0], dv[r][1], part, nb)
            engs[r].sync()
            want = part["flat"].cpu() if want is None else want + part["flat"].cpu()
    np.testing.assert_allclose(acc["flat"].cpu().numpy(), 4 * want.numpy(), rtol=1e-12, atol=1e-9)
    for eng in engs:
        eng.close()


def test_sharded_run_loop_on_one_rank_equals_run_recording(tmp_path):
    """run_recording_sharded without a process group is the drop-in's run loop: calibration on block 0,
    one row per later block, the reference's CSV"""
    from effex_b200.correlator import Correlator
    from effex_b200 import csvio
    S, N, nb = 2**15, 1024, 9
    raw0, raw1 = synth.correlated_pair(nb * S, delay=11, seed=21)
    a = Correlator(num_samp=S, nbins=N, output_file=str(tmp_path / "a.csv"), batch_blocks=4)
    rows_a = a.run_recording(raw0, raw1)
    b = Correlator(num_samp=S, nbins=N, output_file=str(tmp_path / "b.csv"), batch_blocks=4)
    rows_b = sharding.run_recording_sharded(b, raw0, raw1)
    assert a.calibrated_delay == b.calibrated_delay
    np.testing.assert_array_equal(rows_a, rows_b)
    assert open(tmp_path / "a.csv", "rb").read() == open(tmp_path / "b.csv", "rb").read()
    p0, p1 = tmp_path / "c0.iq", tmp_path / "c1.iq"
    raw0.tofile(p0); raw1.tofile(p1)
    c = Correlator(num_samp=S, nbins=N, output_file=str(tmp_path / "c.csv"), batch_blocks=4)
    rows_c = sharding.run_recording_sharded(c, str(p0), str(p1))
    np.testing.assert_array_equal(rows_a, rows_c)
    a.close(); b.close(); c.close()


@pytest.mark.parametrize("world", [2, 4])
def test_torchrun_cli_time_shards_the_run(tmp_path, world):
    """`torchrun --nproc-per-node N -m effex_b200 ...` writes the CSV of the one-GPU command: same header,
    same number of rows, rows equal to float32 rounding (the batches differ, so not bit for bit)"""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import subprocess
    import sys
    from effex_b200 import csvio
    S, N = 2**15, 1024
    nb = int(np.ceil(1 * 2.4e6 / S))                      # --time 1 (the reference refuses less): 74 blocks
    raw0, raw1 = synth.tiled_recording(nb, S, base_blocks=8, delay=11, seed=21)
    p0, p1 = tmp_path / "c0.iq", tmp_path / "c1.iq"
    raw0.tofile(p0); raw1.tofile(p1)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    common = ["--time", "1", "--num_samp", str(S), "--resolution", str(N), "--omit_plot", "1",
              "--loglevel", "ERROR", "--input0", str(p0), "--input1", str(p1)]
    one = subprocess.run([sys.executable, "-m", "effex_b200", *common, "--output", str(tmp_path / "one.csv")],
                         cwd=root, capture_output=True, text=True, timeout=300)
    assert one.returncode == 0, one.stderr[-1500:]
    port = 29700 + os.getpid() % 2000 + world
    many = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                           "--master-addr", "127.0.0.1", "--master-port", str(port), "-m", "effex_b200", *common,
                           "--output", str(tmp_path / "many.csv")], cwd=root, capture_output=True, text=True, timeout=600)
    assert many.returncode == 0, many.stderr[-1500:]
    m1, r1 = csvio.read_rows(str(tmp_path / "one.csv"))
    m2, r2 = csvio.read_rows(str(tmp_path / "many.csv"))
    assert m1 == m2 and r1.shape == r2.shape == (nb - 1, N)
    np.testing.assert_allclose(r2, r1, rtol=0, atol=2e-6 * np.abs(r1).max())
    assert open(tmp_path / "one.csv").readlines()[:2] == open(tmp_path / "many.csv").readlines()[:2]
