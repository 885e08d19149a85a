"""Host-side mirror of effex's `Correlator` for the spectrum-mode hot path.

Same constructor parameters, property validation, private DSP method names and
`.csv` output as the reference (`effex/effex.py`), so its tests and callers can
be pointed here.  What differs, on purpose:

  * blocks come from a recording (uint8 interleaved IQ per channel, host or
    device) instead of two live RTL-SDRs -- the SDR/queue/state-machine
    plumbing of effex.py:76-89, :326-473, :630-664 is out of scope;
  * `gpu_iq_0/1` hold the RAW bytes of the current block pair on the device;
    unpack + DC removal (effex.py:391-395) happen inside the kernels;
  * all arithmetic runs in libeffex_fx.so (hand-written sm_100a CUDA) through
    `FxEngine`; there is no cupy/cuSignal and no CPU fallback.
"""
from __future__ import annotations

import logging
import time

import numpy as np
import torch

from .engine import FxEngine, pfb_window, rot_vector

_STATES = ('OFF', 'STARTUP', 'RUN', 'CALIBRATE', 'SHUTDOWN')
_MODES = ('SPECTRUM', 'CONTINUUM', 'TEST')


class Correlator:
    _states = _STATES
    _modes = _MODES

    def __init__(self, run_time=1, bandwidth=2.4e6, frequency=1.4204e9, num_samp=2**18, nbins=2**12,
                 gain=49.6, mode='SPECTRUM', loglevel='INFO', device=0, extended=False, batch_blocks=64,
                 output_file=None):
        self.logger = logging.getLogger("effex_b200")
        self.logger.setLevel(getattr(logging, loglevel))
        self.device = int(device)
        self.extended = bool(extended)       # lift the [2^8, 2^18] num_samp clamp (BASELINE configs 3, 5)
        self.batch_blocks = int(batch_blocks)

        self.run_time = run_time
        self.bandwidth = bandwidth
        self.frequency = frequency
        self.num_samp = num_samp
        self.nbins = nbins
        self.gain = gain
        self._state = 'OFF'
        self.mode = mode

        self.ntaps = 4                                   # effex.py:115
        S = int(self.num_samp)
        n_int = S // self.ntaps // self.nbins            # effex.py:118-124
        assert n_int >= 1, ('Assertion failed: there must be at least 1 window of length n_branches*ntaps '
                            f'in each input timeseries.\ntimeseries len: {S}\nn_branches: {self.nbins}\n'
                            f'ntaps: {self.ntaps}\nn_branches*ntaps: {self.nbins * self.ntaps}')
        self.window = pfb_window(self.ntaps, self.nbins)  # effex.py:126-127 (float64, host)

        self.calibrated_delay = 0                         # seconds, effex.py:132
        self.output_file = output_file or time.strftime('visibilities_%Y%m%d-%H%M%S') + '.csv'
        crit_delay = 1 / self.frequency                   # effex.py:151-155
        self.test_delay_sweep_step = crit_delay / 2
        self.test_delay_offset = self.test_delay_sweep_step * 1600

        self._engines = {}
        self._rot_key = None
        self.gpu_iq_0 = None      # uint8 CUDA tensors: raw bytes of the current block pair
        self.gpu_iq_1 = None

    # ---- properties: same validation as effex.py:231-320 ----------------------
    @property
    def state(self):
        return self._state

    @property
    def run_time(self):
        return self._run_time

    @run_time.setter
    def run_time(self, run_time):
        if run_time < 1:
            raise ValueError(f'run time {run_time} is not allowed; run times must be >= 1 second.')
        self._run_time = run_time

    @property
    def bandwidth(self):
        return self._bandwidth

    @bandwidth.setter
    def bandwidth(self, value):
        if value > 2.8e6:
            self.logger.warning(f'Bandwidth value {value} is greater than 2800000.0, and RtlSdrs may not be stable.')
        self._bandwidth = value

    @property
    def frequency(self):
        return self._frequency

    @frequency.setter
    def frequency(self, value):
        self._frequency = value

    @property
    def num_samp(self):
        return self._num_samp

    @num_samp.setter
    def num_samp(self, value):
        if not self.extended:                 # effex.py:277-284 (stores the un-rounded value when in range)
            int_val = int(round(value))
            if int_val < 2**8:
                value = 2**8
            elif int_val > 2**18:
                value = 2**18
        self._num_samp = value

    @property
    def nbins(self):
        return self._nbins

    @nbins.setter
    def nbins(self, value):
        self._nbins = value

    @property
    def gain(self):
        return self._gain

    @gain.setter
    def gain(self, value):
        self._gain = value

    @property
    def mode(self):
        return self._mode

    @mode.setter
    def mode(self, input_mode):
        input_mode = input_mode.upper()
        if input_mode in self._modes:
            self._mode = input_mode
        else:
            raise ValueError(f'Mode input {input_mode} is not in known modes: {self._modes}')

    def close(self):
        for e in self._engines.values():
            e.close()
        self._engines.clear()

    # ---- engines ------------------------------------------------------------
    def _engine(self, num_samp, nbins, ntaps, window=None, max_blocks=1) -> FxEngine:
        key = (int(num_samp), int(nbins), int(ntaps))
        eng = self._engines.get(key)
        if eng is None or eng.max_blocks < max_blocks:
            if eng is not None:
                eng.close()
            eng = FxEngine(key[0], key[1], key[2], device=self.device, max_blocks=max(max_blocks, 1),
                           window=window)
            self._engines[key] = eng
            if key == (int(self.num_samp), int(self.nbins), self.ntaps):
                self._rot_key = None
        elif window is not None and (eng.window.shape != np.shape(window) or not np.array_equal(eng.window, window)):
            eng.set_window(window)
        return eng

    def _main_engine(self, max_blocks=1) -> FxEngine:
        eng = self._engine(self.num_samp, self.nbins, self.ntaps, self.window, max_blocks)
        key = (id(eng), self.bandwidth, self.frequency, self.calibrated_delay)
        if key != self._rot_key:      # host float64 rot table, rebuilt only when the delay changes
            eng.set_rot(rot_vector(self.nbins, self.bandwidth, self.frequency, self.calibrated_delay))
            self._rot_key = key
        return eng

    # ---- DSP methods (the reference's "operator API", effex.py:476-627) --------
    def _spectrometer_poly(self, x, ntaps, n_branches, window):
        """effex.py:530-555: x complex (numpy / torch) -> (P, N) complex64 CUDA tensor."""
        n = len(x)
        if n // n_branches < 1:
            raise ValueError("input shorter than one frame")
        return self._engine(n, n_branches, ntaps, np.asarray(_to_numpy(window), dtype=np.float64)).pfb(x)

    def _pfb_xcorr(self):
        """effex.py:497-527 on the raw block pair in gpu_iq_0/1."""
        eng = self._main_engine()
        xspec = eng.process(self.gpu_iq_0, self.gpu_iq_1, 1)[0]
        if self.mode in ['CONTINUUM', 'TEST']:
            return xspec.mean() / self.bandwidth
        return xspec

    def _run_task(self):
        return self._pfb_xcorr()

    def _calibrate_task(self):
        self.calibrated_delay = self._estimate_delay(self.gpu_iq_0, self.gpu_iq_1, self.bandwidth)
        self.logger.info('Estimated delay (us): {}'.format(1e6 * self.calibrated_delay))

    def _estimate_delay(self, iq_0, iq_1, rate):
        total_delay = self._estimate_delay_gaussian(iq_0, iq_1, rate)
        if self.mode in ['TEST']:
            total_delay -= self.test_delay_offset
        return total_delay

    def _estimate_delay_gaussian(self, iq_0, iq_1, rate, n_blocks=1):
        """effex.py:583-627.  iq_k: complex arrays (as the reference's tests pass) or
        raw uint8 CUDA tensors (as the run loop holds them)."""
        raw = isinstance(iq_0, torch.Tensor) and iq_0.dtype == torch.uint8
        n = (iq_0.numel() // 2 if raw else len(iq_0)) // n_blocks
        if not raw:
            assert len(iq_0) == len(iq_1), ('Algorithm assumes input complex timeseries'
                                            + ' are of equal length.')
        eng = self._engine(n, _lag_nbins(n), 1, None, max_blocks=n_blocks)
        n, imax, xprev, xbest, xnext = eng.lag(iq_0, iq_1, n_blocks)
        if xprev < 0 or xnext < 0:
            raise IndexError('correlation peak at the edge of the lag window')   # reference: TODO at :619
        delta_subpixel = 0.5 * (np.log(xprev) - np.log(xnext)) / (
            np.log(xprev) - 2. * np.log(xbest) + np.log(xnext))
        return (n - (imax + delta_subpixel)) / rate

    # ---- CSV (effex.py:667-693) -------------------------------------------------
    def _write_metadata(self):
        from .csvio import write_metadata
        write_metadata(self.output_file, self.run_time, self.bandwidth, self.frequency, self.num_samp,
                       self.nbins, self.gain, self.mode)

    def _write_data(self, rows):
        from .csvio import append_rows
        append_rows(self.output_file, rows)

    def _start_writer(self):
        """The reference formats and writes rows in a thread of its own (effex.py:457-460, :687-696) so that
        the loop never waits for the disk.  Same here: batches of rows go through a queue to one writer
        thread (row order = queue order); formatting runs in the library (fx_csv_format_rows, ctypes drops
        the GIL), so batch k is formatted and written while batch k+1 is on the GPU."""
        from .csvio import RowWriter
        return RowWriter(self.output_file)

    # ---- run over a recording ------------------------------------------------------
    def run_recording(self, raw0, raw1, write_csv=True, calibrate=True):
        """Process a two-channel recording (uint8 interleaved IQ, numpy).  Mirrors one
        pass of run_state_machine (effex.py:326-417): the first block pair is used for
        delay calibration and produces no row (:399-401); every later block pair gives
        one row.  TEST mode advances the delay per block (:403-404) and therefore runs
        block by block; SPECTRUM/CONTINUUM go through the batched host pipeline."""
        S = int(self.num_samp)
        n_blocks = min(raw0.size, raw1.size) // (2 * S)
        if write_csv:
            self._write_metadata()
        first = 0
        if calibrate and n_blocks > 0:
            self.gpu_iq_0 = torch.from_numpy(raw0[:2 * S]).to(f"cuda:{self.device}")
            self.gpu_iq_1 = torch.from_numpy(raw1[:2 * S]).to(f"cuda:{self.device}")
            self._calibrate_task()
            first = 1
        rows = []
        if self.mode == 'TEST':
            for b in range(first, n_blocks):
                self.gpu_iq_0 = torch.from_numpy(raw0[2 * S * b:2 * S * (b + 1)]).to(f"cuda:{self.device}")
                self.gpu_iq_1 = torch.from_numpy(raw1[2 * S * b:2 * S * (b + 1)]).to(f"cuda:{self.device}")
                self.calibrated_delay += self.test_delay_sweep_step
                rows.append(np.array([complex(self._run_task().item())]))
                if write_csv:
                    self._write_data(rows[-1:])
            return np.array(rows).reshape(-1)
        nb = n_blocks - first
        if nb <= 0:
            return np.zeros((0, self.nbins), dtype=np.complex64)
        eng = self._main_engine(max_blocks=min(self.batch_blocks, nb))
        out = np.empty((nb, self.nbins), dtype=np.complex64)
        writer = self._start_writer() if write_csv else None
        try:
            for b0 in range(0, nb, self.batch_blocks):
                n = min(self.batch_blocks, nb - b0)
                lo = 2 * S * (first + b0)
                eng.process_host(raw0[lo:lo + 2 * S * n], raw1[lo:lo + 2 * S * n], n, out=out[b0:b0 + n])
                chunk = out[b0:b0 + n]
                if self.mode == 'CONTINUUM':
                    chunk = (chunk.astype(np.complex128).mean(axis=1) / self.bandwidth).reshape(-1, 1)
                if writer:
                    writer.put(chunk)              # formatted + written while the next batch computes
        finally:
            if writer:
                writer.close()
        if self.mode == 'CONTINUUM':
            return out.astype(np.complex128).mean(axis=1) / self.bandwidth
        return out


def run_files(cor: "Correlator", path0, path1, write_csv: bool = True, calibrate: bool = True,
              max_blocks: int | None = None):
    """Spectrum/continuum run over two raw byte sources streamed batch by batch through the host pipeline:
    regular files (`ingest.RecordingReader`) or FIFOs / pipes / file objects fed by e.g. two `rtl_sdr`
    processes (`ingest.StreamReader`; the reference's `_streaming` producers, effex.py:630-664).  Same row
    semantics as run_recording: the first block pair calibrates the delay and yields no row (:399-401)."""
    from .ingest import open_reader, _is_regular_file
    S = int(cor.num_samp)
    if cor.mode == 'TEST':
        raise ValueError("TEST mode sweeps the delay per block; use run_recording")
    if write_csv:
        cor._write_metadata()
    seekable = _is_regular_file(path0) and _is_regular_file(path1)
    skip = 0
    if calibrate and seekable:
        head0 = np.fromfile(path0, dtype=np.uint8, count=2 * S)
        head1 = np.fromfile(path1, dtype=np.uint8, count=2 * S)
        if head0.size == 2 * S and head1.size == 2 * S:
            cor.gpu_iq_0 = torch.from_numpy(head0).to(f"cuda:{cor.device}")
            cor.gpu_iq_1 = torch.from_numpy(head1).to(f"cuda:{cor.device}")
            cor._calibrate_task()
            skip = 1
    reader = open_reader(path0, path1, S, batch_blocks=cor.batch_blocks, skip_blocks=skip,
                         max_blocks=max_blocks if seekable or max_blocks is None else max_blocks)
    eng = None
    rows = []
    writer = cor._start_writer() if write_csv else None
    pending_cal = calibrate and not seekable          # a stream calibrates on the first block it delivers
    try:
        for raw0, raw1, first, nb in reader:
            if pending_cal:
                cor.gpu_iq_0 = torch.from_numpy(raw0[:2 * S]).to(f"cuda:{cor.device}")
                cor.gpu_iq_1 = torch.from_numpy(raw1[:2 * S]).to(f"cuda:{cor.device}")
                cor._calibrate_task()
                pending_cal = False
                raw0, raw1, nb = raw0[2 * S:], raw1[2 * S:], nb - 1
                if nb == 0:
                    continue
            if eng is None:      # after the calibration: the rot table is built from the calibrated delay
                eng = cor._main_engine(max_blocks=max(1, cor.batch_blocks))
            out = eng.process_host(raw0, raw1, nb)
            if cor.mode == 'CONTINUUM':
                out = (out.astype(np.complex128).mean(axis=1) / cor.bandwidth).reshape(-1, 1)
            if writer:
                writer.put(out)
            rows.append(out)
    finally:
        if writer:
            writer.close()
    if not rows:
        return np.zeros((0, cor.nbins), dtype=np.complex64)
    out = np.concatenate(rows, axis=0)
    return out.reshape(-1) if cor.mode == 'CONTINUUM' else out


def _to_numpy(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def _lag_nbins(n: int) -> int:
    """The lag search does not use the PFB; any legal nbins <= n will do for the handle."""
    nb = 8
    while nb * 2 <= min(n, 4096):
        nb *= 2
    return nb
