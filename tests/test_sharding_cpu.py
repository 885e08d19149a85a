"""world_size-2 gloo test of the N>1 host logic (shard ranges, the one reduce,
the ordered gather) with the oracle standing in for the GPU engine."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from effex_b200 import sharding, synth
from oracle import fx_oracle as orc

S, N, T, NB = 2048, 128, 4, 7
BW, FC = 2.4e6, 1.4204e9


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 550, 32959):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s1 == s0 + c0
            assert spans[-1][0] + spans[-1][1] == n
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _oracle_compute(raw0, raw1):
    w = orc.pfb_window(T, N)

    def compute(start, count):
        rows, ax, a0, a1 = [], np.zeros(N, complex), np.zeros(N), np.zeros(N)
        for b in range(start, start + count):
            sl = slice(2 * S * b, 2 * S * (b + 1))
            f0 = orc.spectrometer_poly(orc.block_from_u8(raw0[sl]), T, N, w)
            f1 = orc.spectrometer_poly(orc.block_from_u8(raw1[sl]), T, N, w)
            x = f0 * np.conj(f1)
            rows.append(np.fft.fftshift(x.mean(axis=0)))
            ax += x.sum(axis=0); a0 += (abs(f0) ** 2).sum(axis=0); a1 += (abs(f1) ** 2).sum(axis=0)
        acc = {"x": torch.from_numpy(ax.view(np.float64).copy()), "a0": torch.from_numpy(a0),
               "a1": torch.from_numpy(a1), "frames": torch.tensor([float(count * (S // N))], dtype=torch.float64)}
        r = np.array(rows, dtype=np.complex64).reshape(count, N)
        return torch.from_numpy(r), acc
    return compute


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    raw0, raw1 = synth.correlated_pair(NB * S, delay=3, seed=9)
    rows, acc = sharding.sharded_run(_oracle_compute(raw0, raw1), NB)
    if rank == 0:
        x, a0, a1 = sharding.finish_integration(acc)
        np.savez(out, rows=rows.numpy(), x=x, a0=a0, a1=a1, frames=acc["frames"].numpy())
    else:
        assert rows is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_run_matches_single_process(tmp_path, world):
    out = str(tmp_path / "r.npz")
    port = 29500 + os.getpid() % 2000 + world
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = np.load(out)
    raw0, raw1 = synth.correlated_pair(NB * S, delay=3, seed=9)
    rows1, acc1 = _oracle_compute(raw0, raw1)(0, NB)
    x1, a01, a11 = sharding.finish_integration(acc1)
    np.testing.assert_array_equal(got["rows"], rows1.numpy())            # rows in block order, untouched
    assert got["frames"][0] == NB * (S // N)
    np.testing.assert_allclose(got["x"], x1, rtol=1e-12, atol=1e-18)     # reduce == single pass
    np.testing.assert_allclose(got["a0"], a01, rtol=1e-12)
    np.testing.assert_allclose(got["a1"], a11, rtol=1e-12)
    ref = orc.process_recording_u8(raw0, raw1, S, N, BW, FC, 0.0, T, 0, NB)
    np.testing.assert_allclose(got["x"], ref.mean(axis=0), rtol=1e-10, atol=1e-16)


# ---- streaming-history sharding (halo + global byte sums + one reduce) with an oracle-backed engine ----
class _OracleEngine:
    """Stands in for FxEngine on CPU: same methods stream_integrate() uses, float64 oracle inside."""
    num_samp, nbins, ntaps = S, N, T

    def span_sums(self, a, b, nb):
        a, b = a.numpy(), b.numpy()
        return np.array([a[0::2].sum(dtype=np.uint64), a[1::2].sum(dtype=np.uint64),
                         b[0::2].sum(dtype=np.uint64), b[1::2].sum(dtype=np.uint64)], dtype=np.uint64)

    def new_accumulators(self):
        flat = torch.zeros(4 * N + 1, dtype=torch.float64)
        return {"flat": flat, "x": flat[:2 * N], "a0": flat[2 * N:3 * N], "a1": flat[3 * N:4 * N], "frames": flat[4 * N:]}

    def integrate_stream(self, a, b, acc, nb, halo0, halo1, sums, total):
        w = orc.pfb_window(T, N)
        hist = (T - 1) if halo0 is not None else 0

        def chan(raw, halo, si, sq):
            raw = np.concatenate([halo.numpy(), raw.numpy()]) if halo is not None else raw.numpy()
            x = (raw[0::2].astype(np.float64) - float(si) / total) / 127.5 \
                + 1j * (raw[1::2].astype(np.float64) - float(sq) / total) / 127.5
            return orc.spectrometer_poly(x, T, N, w)[hist:]          # frames of the halo itself are not output
        f0 = chan(a, halo0, sums[0], sums[1])
        f1 = chan(b, halo1, sums[2], sums[3])
        acc["x"] += torch.from_numpy((f0 * np.conj(f1)).sum(axis=0).view(np.float64).copy())
        acc["a0"] += torch.from_numpy((abs(f0) ** 2).sum(axis=0))
        acc["a1"] += torch.from_numpy((abs(f1) ** 2).sum(axis=0))
        acc["frames"] += f0.shape[0]
        return acc


def _stream_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    raw0, raw1 = synth.correlated_pair(NB * S, delay=3, dc0=0.03, dc1=-0.02j, seed=9)
    start, count = sharding.shard_range(NB, world, rank)
    sl = slice(2 * S * start, 2 * S * (start + count))
    acc = sharding.stream_integrate(_OracleEngine(), torch.from_numpy(raw0[sl].copy()), torch.from_numpy(raw1[sl].copy()))
    if rank == 0:
        x, a0, a1 = sharding.finish_integration(acc)
        np.savez(out, x=x, a0=a0, frames=acc["frames"].numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_stream_integrate_across_ranks_equals_one_giant_block(tmp_path):
    out = str(tmp_path / "s.npz")
    port = 31500 + os.getpid() % 2000
    mp.spawn(_stream_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    raw0, raw1 = synth.correlated_pair(NB * S, delay=3, dc0=0.03, dc1=-0.02j, seed=9)
    w = orc.pfb_window(T, N)
    f0 = orc.spectrometer_poly(orc.block_from_u8(raw0), T, N, w)       # the recording as ONE block
    f1 = orc.spectrometer_poly(orc.block_from_u8(raw1), T, N, w)
    assert got["frames"][0] == f0.shape[0]
    np.testing.assert_allclose(got["x"], np.fft.fftshift((f0 * np.conj(f1)).mean(axis=0)), rtol=1e-9, atol=1e-16)
    np.testing.assert_allclose(got["a0"], np.fft.fftshift((abs(f0) ** 2).mean(axis=0)), rtol=1e-9)
