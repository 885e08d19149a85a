"""Command line of the drop-in: the reference's flags (effex.py:703-770), same
names, defaults and clamps, plus where the two channels' raw bytes come from
(the reference reads two live RTL-SDRs; here: rtl_sdr-style uint8 IQ files, or
a synthetic correlated-noise recording).

    python -m effex_b200 --time 60 --bandwidth 2.4e6 --frequency 1.4204e9 \\
        --num_samp 262144 --resolution 4096 --mode spectrum --input0 a.iq --input1 b.iq
"""
from __future__ import annotations

import argparse
import os

import numpy as np


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description='B200-native FX correlator hot path for effex.',
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('--time', '-T', default=1, type=float, dest='run_time',
                        help='(sec) Total amount of time to run correlator.')
    parser.add_argument('--bandwidth', '-B', default=2.4e6, type=float, dest='bandwidth',
                        help='(Hz) Receiver bandwidth = complex sample rate. Applied to both channels.')
    parser.add_argument('--frequency', '-F', default=1.4204e9, type=float, dest='fc',
                        help='(Hz) Receiver center tuning frequency. Applied to both channels.')
    parser.add_argument('--num_samp', '-N', default=2**18, type=int, dest='num_samp',
                        help='(int) Number of samples per block (clamped to [2^8, 2^18] unless --extended).')
    parser.add_argument('--resolution', '-R', default=2**12, type=int, dest='nfft',
                        help='(int) Number of FFT bins to use in processing and plotting.')
    parser.add_argument('--gain', '-G', default=49.6, type=float, dest='gain',
                        help='(dB) Tuner gain; recorded in the file header only.')
    parser.add_argument('--mode', '-M', default='spectrum', type=str, choices=['continuum', 'spectrum', 'test'],
                        dest='mode', help='(str) continuum | spectrum | test.')
    parser.add_argument('--omit_plot', '-P', default=False, type=bool, dest='omit_plot',
                        help='If True, skip the matplotlib post-processing step (any non-empty string is True, '
                             'as in the reference).')
    parser.add_argument('--loglevel', '-L', default='INFO', type=str,
                        choices=['INFO', 'WARNING', 'DEBUG', 'ERROR', 'CRITICAL'], dest='loglevel')
    # --- not in the reference -------------------------------------------------
    parser.add_argument('--input0', default=None,
                        help='raw uint8 interleaved IQ, channel 0: a file, or a FIFO fed by e.g. `rtl_sdr -d 0 -`')
    parser.add_argument('--input1', default=None, help='raw uint8 interleaved IQ, channel 1 (file or FIFO)')
    parser.add_argument('--synthetic-delay', default=37, type=int,
                        help='without input files: synthetic correlated noise, channel 1 lagging by this many samples')
    parser.add_argument('--extended', action='store_true',
                        help='lift the reference clamp num_samp <= 2^18 (BASELINE configs 3 and 5)')
    parser.add_argument('--device', default=0, type=int)
    parser.add_argument('--output', default=None, help='csv path (default visibilities_%%Y%%m%%d-%%H%%M%%S.csv)')
    parser.add_argument('--timing', action='store_true', help='print a JSON line with the wall time of each phase')
    return parser


def main(argv=None):
    import time
    t_start = time.perf_counter()
    args = build_parser().parse_args(argv)
    from .correlator import Correlator
    from . import synth, csvio
    phases = {"import_s": time.perf_counter() - t_start}
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        args.device = int(os.environ.get("LOCAL_RANK", "0"))
    cor = Correlator(run_time=args.run_time, bandwidth=args.bandwidth, frequency=args.fc, num_samp=args.num_samp,
                     nbins=args.nfft, gain=args.gain, mode=args.mode, loglevel=args.loglevel, device=args.device,
                     extended=args.extended, output_file=args.output)
    S = int(cor.num_samp)
    n_blocks = int(np.ceil(cor.run_time * cor.bandwidth / S))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # launched with torchrun: one process per GPU, the blocks of the run time-sharded over the ranks
        return _main_sharded(args, cor, n_blocks, world, phases, t_start)
    if args.input0 and args.input1 and args.mode != 'test':
        from .correlator import run_files
        # streamed in chunks of whole blocks; a FIFO / pipe ends at EOF or after --time worth of blocks
        run_files(cor, args.input0, args.input1, max_blocks=n_blocks)
    else:
        if args.input0 and args.input1:
            raw0 = np.fromfile(args.input0, dtype=np.uint8, count=2 * S * n_blocks)
            raw1 = np.fromfile(args.input1, dtype=np.uint8, count=2 * S * n_blocks)
        else:
            # synthetic recording: 8 fresh blocks repeated, presented as views (an hour of data is 17 GB per channel)
            raw0, raw1 = synth.tiled_recording_lazy(n_blocks, S, base_blocks=min(8, n_blocks), delay=args.synthetic_delay,
                                                    window_blocks=cor.batch_blocks)
        phases["input_s"] = time.perf_counter() - t_start - phases["import_s"]
        t_run = time.perf_counter()
        cor.run_recording(raw0, raw1)
        phases["run_s"] = time.perf_counter() - t_run
    cor.close()
    print(f'wrote {cor.output_file}; estimated delay {1e6 * cor.calibrated_delay:.6f} us')
    if args.timing:
        import json
        phases["total_s"] = time.perf_counter() - t_start
        phases["csv_bytes"] = os.path.getsize(cor.output_file)
        phases["rows"] = max(n_blocks - 1, 0)
        print(json.dumps({"effex_b200_cli_timing": phases}))
    if not args.omit_plot:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            print('matplotlib is not installed: skipping the plot (the reference\'s post_process only plots)')
            return 0
        from .post_process import post_process
        meta, rows = csvio.read_rows(cor.output_file)
        post_process(rows, args.bandwidth, args.fc, args.nfft, args.mode, args.omit_plot,
                     test_delay_sweep_step=cor.test_delay_sweep_step if args.mode == 'test' else 0)
    return 0


def _main_sharded(args, cor, n_blocks, world, phases, t_start):
    """`torchrun --nproc-per-node N -m effex_b200 ...`: rank 0 calibrates and writes the CSV, every rank
    correlates its own contiguous range of blocks (north_star: time blocks across the GPUs of the box)."""
    import time
    import torch
    import torch.distributed as dist
    from . import sharding, synth
    if args.mode == 'test':
        raise SystemExit("--mode test sweeps the delay block by block and runs on one GPU only")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(cor.device)
    dist.init_process_group("nccl", device_id=torch.device("cuda", cor.device))
    rank = dist.get_rank()
    S = int(cor.num_samp)
    if args.output is None:                      # every rank must agree on the file name rank 0 writes
        name = [cor.output_file]
        dist.broadcast_object_list(name, src=0)
        cor.output_file = name[0]
    if args.input0 and args.input1:
        src0, src1 = args.input0, args.input1
    else:
        src0, src1 = synth.tiled_recording_lazy(n_blocks, S, base_blocks=min(8, n_blocks), delay=args.synthetic_delay,
                                                window_blocks=cor.batch_blocks)
    t_run = time.perf_counter()
    rows = sharding.run_recording_sharded(cor, src0, src1, n_blocks)
    dist.barrier()
    if rank == 0:
        print(f'wrote {cor.output_file} ({len(rows)} rows from {world} GPUs); estimated delay {1e6 * cor.calibrated_delay:.6f} us')
        if args.timing:
            import json
            phases["run_s"] = time.perf_counter() - t_run
            phases["total_s"] = time.perf_counter() - t_start
            phases["csv_bytes"] = os.path.getsize(cor.output_file)
            phases["rows"] = len(rows)
            phases["gpus"] = world
            print(json.dumps({"effex_b200_cli_timing": phases}))
    cor.close()
    dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
