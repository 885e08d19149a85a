"""FxEngine: thin Python face of the C ABI (one handle = one GPU, one shape).

PyTorch is used here only for device memory and stream ordering; every kernel
that runs is ours, launched by libeffex_fx.so through ctypes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.signal
import torch

from . import _lib


class FxError(RuntimeError):
    pass


class FxCommError(FxError):
    """cross-GPU reduce failed (no peer access, or a rank never arrived)"""


def pfb_window(ntaps: int, nbins: int) -> np.ndarray:
    """Prototype filter of effex.py:126-127 (float64, host)."""
    L = int(ntaps) * int(nbins)
    return (scipy.signal.get_window("hamming", L)
            * scipy.signal.firwin(L, cutoff=1.0 / nbins, window="rectangular"))


def rot_vector(nbins: int, bandwidth: float, frequency: float, delay: float) -> np.ndarray:
    """rot[c] of effex.py:516-519 in float64, phase reduced mod 1 cycle before
    exp so the float32 copy on the device keeps ~1e-7 accuracy."""
    freqs = np.fft.fftfreq(nbins, d=1.0 / bandwidth) + frequency
    cycles = np.mod(freqs * delay, 1.0)
    return np.exp(2j * np.pi * cycles)


def _raise(lib, handle, rc, what):
    msg = lib.fx_last_error(handle)
    msg = msg.decode() if msg else ""
    text = f"{what} failed ({rc}): {msg}"
    if rc in (_lib.FX_ERR_INVALID, _lib.FX_ERR_UNSUPPORTED):
        raise ValueError(text)
    if rc == _lib.FX_ERR_COMM:
        raise FxCommError(text)
    raise FxError(text)


class FxEngine:
    def __init__(self, num_samp: int, nbins: int, ntaps: int = 4, device: int = 0, max_blocks: int = 1,
                 dc_remove: bool = True, force_generic: bool = False, window: np.ndarray | None = None,
                 lockstep_kernel: bool = False, cross_only: bool = False):
        self.lib = _lib.load()
        self.num_samp, self.nbins, self.ntaps = int(num_samp), int(nbins), int(ntaps)
        self.device = int(device)
        self.max_blocks = int(max_blocks)
        cfg = _lib.FxConfig(self.device, self.ntaps, self.nbins, 1 if dc_remove else 0, self.num_samp,
                            self.max_blocks, (_lib.FX_FLAG_FORCE_GENERIC if force_generic else 0)
                            | (_lib.FX_FLAG_LOCKSTEP_KERNEL if lockstep_kernel else 0)
                            | (_lib.FX_FLAG_CROSS_ONLY if cross_only else 0))
        h = C.c_void_p()
        rc = self.lib.fx_create(C.byref(cfg), C.byref(h))
        if rc != _lib.FX_OK:
            _raise(self.lib, None, rc, "fx_create")
        self.h = h
        self.frames_per_block = self.num_samp // self.nbins
        self.tdev = torch.device("cuda", self.device)
        self.stream = torch.cuda.ExternalStream(self.lib.fx_stream(self.h), device=self.tdev)
        self.stream_aux = torch.cuda.ExternalStream(self.lib.fx_stream_aux(self.h), device=self.tdev)
        self.set_window(pfb_window(self.ntaps, self.nbins) if window is None else window)

    # ---- lifetime ---------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.fx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != _lib.FX_OK:
            _raise(self.lib, self.h, rc, what)

    @property
    def fused(self) -> bool:
        return bool(self.lib.fx_uses_fused(self.h))

    def sync(self):
        self._check(self.lib.fx_sync(self.h), "fx_sync")

    # ---- parameters -------------------------------------------------------
    def set_window(self, window):
        w = np.ascontiguousarray(np.asarray(window, dtype=np.float64))
        self._check(self.lib.fx_set_taps(self.h, w.ctypes.data_as(C.POINTER(C.c_double)), w.size), "fx_set_taps")
        self.window = w

    def set_rot(self, rot):
        if rot is None:
            self._check(self.lib.fx_set_rot(self.h, None, 0), "fx_set_rot")
            return
        r = np.ascontiguousarray(np.asarray(rot, dtype=np.complex128)).view(np.float64)
        self._check(self.lib.fx_set_rot(self.h, r.ctypes.data_as(C.POINTER(C.c_double)), r.size // 2), "fx_set_rot")

    def set_delay(self, bandwidth: float, frequency: float, delay: float):
        self.set_rot(rot_vector(self.nbins, bandwidth, frequency, delay))

    # ---- stream ordering with torch ----------------------------------------
    def _enter(self, inputs_ready: bool = False):
        """Order the engine's streams after the caller's current stream.  `inputs_ready=True`
        (the caller guarantees the input tensors are complete, e.g. a resident recording) skips it,
        which lets the byte-sum pre-pass of this call overlap the fused kernel of the previous one."""
        if inputs_ready:
            return
        cur = torch.cuda.current_stream(self.tdev)
        self.stream.wait_stream(cur)
        self.stream_aux.wait_stream(cur)      # the byte-sum pre-pass reads the inputs on this stream

    def _exit(self):
        torch.cuda.current_stream(self.tdev).wait_stream(self.stream)

    def _raw(self, t: torch.Tensor, n_blocks: int) -> int:
        if t.dtype != torch.uint8 or not t.is_cuda or t.device.index != self.device or not t.is_contiguous():
            raise ValueError("raw IQ must be a contiguous uint8 CUDA tensor on the engine's device")
        if t.numel() < 2 * self.num_samp * n_blocks:
            raise ValueError("raw IQ tensor is shorter than n_blocks * 2 * num_samp bytes")
        return t.data_ptr()

    # ---- hot path -----------------------------------------------------------
    def process(self, iq0: torch.Tensor, iq1: torch.Tensor, n_blocks: int | None = None, autos: bool = False,
                out=None, acc=None, inputs_ready: bool = False):
        """fx_process: one fftshifted, rot-applied cross-spectrum row per block.
        With `acc` (see new_accumulators) the same kernel run also adds the
        un-normalised sums into the float64 accumulators (fx_process_acc)."""
        if n_blocks is None:
            n_blocks = iq0.numel() // (2 * self.num_samp)
        p0, p1 = self._raw(iq0, n_blocks), self._raw(iq1, n_blocks)
        if out is None:
            x = torch.empty((n_blocks, self.nbins), dtype=torch.complex64, device=self.tdev)
            a0 = torch.empty((n_blocks, self.nbins), dtype=torch.float32, device=self.tdev) if autos else None
            a1 = torch.empty((n_blocks, self.nbins), dtype=torch.float32, device=self.tdev) if autos else None
        else:
            x, a0, a1 = out
        self._enter(inputs_ready)
        pa0 = a0.data_ptr() if a0 is not None else None
        pa1 = a1.data_ptr() if a1 is not None else None
        if acc is None:
            rc = self.lib.fx_process(self.h, p0, p1, n_blocks, x.data_ptr(), pa0, pa1)
        else:
            rc = self.lib.fx_process_acc(self.h, p0, p1, n_blocks, x.data_ptr(), pa0, pa1, acc["x"].data_ptr(),
                                         acc["a0"].data_ptr(), acc["a1"].data_ptr(), acc["frames"].data_ptr())
        self._check(rc, "fx_process")
        self._exit()
        return (x, a0, a1) if autos else x

    # ---- the cross-GPU reduce, owned by the library (fx_comm_*) ---------------------------------
    def comm_export(self, world: int, slot_bytes: int = 0) -> bytes:
        """Create this engine's mailbox; returns the token the peers need (exchange it with any transport)."""
        tok = C.create_string_buffer(_lib.FX_COMM_TOKEN_BYTES)
        self._check(self.lib.fx_comm_export(self.h, int(world), int(slot_bytes), tok), "fx_comm_export")
        return tok.raw

    def comm_attach(self, rank: int, world: int, tokens) -> None:
        blob = b"".join(tokens)
        if len(blob) != world * _lib.FX_COMM_TOKEN_BYTES:
            raise ValueError("need one token per rank")
        self._check(self.lib.fx_comm_attach(self.h, int(rank), int(world), blob), "fx_comm_attach")
        self.comm_rank, self.comm_world = int(rank), int(world)

    @property
    def comm_attached(self) -> bool:
        return getattr(self, "comm_world", 0) > 0

    def comm_fence(self):
        self._check(self.lib.fx_comm_fence(self.h), "fx_comm_fence")

    def process_reduce(self, iq0, iq1, n_blocks: int | None = None, out=None, acc=None, root: int = 0,
                       inputs_ready: bool = False):
        """fx_process_reduce: rows of this rank's blocks + this call's sums ADDED into the root's
        accumulators (acc["flat"] on the root; ignored elsewhere).  Collective over the attached world."""
        if n_blocks is None:
            n_blocks = iq0.numel() // (2 * self.num_samp)
        p0, p1 = self._raw(iq0, n_blocks), self._raw(iq1, n_blocks)
        if out is None:
            x = torch.empty((n_blocks, self.nbins), dtype=torch.complex64, device=self.tdev)
            a0 = a1 = None
        else:
            x, a0, a1 = out
        self._enter(inputs_ready)
        rc = self.lib.fx_process_reduce(self.h, p0, p1, n_blocks, x.data_ptr() if x is not None else None,
                                        a0.data_ptr() if a0 is not None else None,
                                        a1.data_ptr() if a1 is not None else None, int(root),
                                        acc["flat"].data_ptr() if acc is not None else None)
        self._check(rc, "fx_process_reduce")
        self._exit()
        return x

    def integrate_stream_reduce(self, iq0, iq1, acc, n_blocks: int | None = None, halo0=None, halo1=None, sums=None,
                                total_samp: int | None = None, root: int = 0):
        """fx_integrate_stream_reduce: integrate_stream whose sums go to the root's accumulators."""
        if n_blocks is None:
            n_blocks = iq0.numel() // (2 * self.num_samp)
        p0, p1 = self._raw(iq0, n_blocks), self._raw(iq1, n_blocks)
        ph0, ph1 = self._halo_ptrs(halo0, halo1)
        csums = (C.c_uint64 * 4)(*[int(v) for v in sums]) if sums is not None else None
        self._enter()
        rc = self.lib.fx_integrate_stream_reduce(self.h, p0, p1, n_blocks, ph0, ph1, csums,
                                                 int(total_samp) if total_samp else 0, int(root),
                                                 acc["flat"].data_ptr() if acc is not None else None)
        self._check(rc, "fx_integrate_stream_reduce")
        self._exit()
        return acc

    def reduce_inplace(self, buf: torch.Tensor, root: int = 0):
        """fx_reduce_f64 / fx_reduce_f32: buf (float64, float32 or complex64) summed over the ranks into
        the root's buf.  Collective."""
        if not buf.is_cuda or not buf.is_contiguous():
            raise ValueError("reduce_inplace needs a contiguous CUDA tensor")
        self._enter()
        if buf.dtype == torch.float64:
            rc = self.lib.fx_reduce_f64(self.h, buf.data_ptr(), buf.numel(), int(root))
        elif buf.dtype == torch.float32:
            rc = self.lib.fx_reduce_f32(self.h, buf.data_ptr(), buf.numel(), int(root))
        elif buf.dtype == torch.complex64:
            rc = self.lib.fx_reduce_f32(self.h, buf.data_ptr(), 2 * buf.numel(), int(root))
        else:
            raise ValueError("unsupported dtype")
        self._check(rc, "fx_reduce")
        self._exit()
        return buf

    def _halo_ptrs(self, halo0, halo1):
        hb = 2 * (self.ntaps - 1) * self.nbins
        if halo0 is None:
            return None, None
        for t in (halo0, halo1):
            if t is None or t.dtype != torch.uint8 or not t.is_cuda or not t.is_contiguous() or t.numel() != hb:
                raise ValueError(f"halo must be a contiguous uint8 CUDA tensor of {hb} bytes")
        return halo0.data_ptr(), halo1.data_ptr()

    def new_accumulators(self):
        """float64 accumulators of one integration: views into ONE flat buffer, so the cross-GPU
        reduce is a single collective on `acc["flat"]` and clearing is a single memset."""
        n = self.nbins
        flat = torch.zeros(4 * n + 1, dtype=torch.float64, device=self.tdev)
        return {"flat": flat, "x": flat[:2 * n], "a0": flat[2 * n:3 * n], "a1": flat[3 * n:4 * n],
                "frames": flat[4 * n:]}

    def integrate(self, iq0, iq1, acc, n_blocks: int | None = None):
        """fx_integrate: add this call's un-normalised sums into float64 accumulators."""
        if n_blocks is None:
            n_blocks = iq0.numel() // (2 * self.num_samp)
        p0, p1 = self._raw(iq0, n_blocks), self._raw(iq1, n_blocks)
        self._enter()
        rc = self.lib.fx_integrate(self.h, p0, p1, n_blocks, acc["x"].data_ptr(), acc["a0"].data_ptr(),
                                   acc["a1"].data_ptr(), acc["frames"].data_ptr())
        self._check(rc, "fx_integrate")
        self._exit()
        return acc

    def span_sums(self, iq0, iq1, n_blocks: int | None = None) -> np.ndarray:
        """fx_span_sums: exact byte sums {I0, Q0, I1, Q1} of the span (uint64[4])."""
        if n_blocks is None:
            n_blocks = iq0.numel() // (2 * self.num_samp)
        p0, p1 = self._raw(iq0, n_blocks), self._raw(iq1, n_blocks)
        out = (C.c_uint64 * 4)()
        self._enter()
        self._check(self.lib.fx_span_sums(self.h, p0, p1, n_blocks, out), "fx_span_sums")
        return np.array(list(out), dtype=np.uint64)

    def integrate_stream(self, iq0, iq1, acc, n_blocks: int | None = None, halo0=None, halo1=None, sums=None,
                         total_samp: int | None = None):
        """fx_integrate_stream: the span is a piece of one long recording (PFB history carried, `halo`
        = the (ntaps-1)*nbins samples before it as raw bytes, `sums`/`total_samp` = recording-wide byte sums)."""
        if n_blocks is None:
            n_blocks = iq0.numel() // (2 * self.num_samp)
        p0, p1 = self._raw(iq0, n_blocks), self._raw(iq1, n_blocks)
        ph0, ph1 = self._halo_ptrs(halo0, halo1)
        csums = None
        if sums is not None:
            csums = (C.c_uint64 * 4)(*[int(v) for v in sums])
        self._enter()
        rc = self.lib.fx_integrate_stream(self.h, p0, p1, n_blocks, ph0, ph1, csums,
                                          int(total_samp) if total_samp else 0, acc["x"].data_ptr(),
                                          acc["a0"].data_ptr(), acc["a1"].data_ptr(), acc["frames"].data_ptr())
        self._check(rc, "fx_integrate_stream")
        self._exit()
        return acc

    @staticmethod
    def finish_integration(acc, rot=None):
        """Host epilogue of an integration (after any cross-GPU reduce):
        1/frames, conj(rot), fftshift -- effex.py:519-521 in float64."""
        frames = float(acc["frames"].item())
        x = acc["x"].cpu().numpy().view(np.complex128) / frames
        if rot is not None:
            x = x * np.conj(rot)
        a0 = acc["a0"].cpu().numpy() / frames
        a1 = acc["a1"].cpu().numpy() / frames
        return np.fft.fftshift(x), np.fft.fftshift(a0), np.fft.fftshift(a1)

    def process_host(self, raw0: np.ndarray, raw1: np.ndarray, n_blocks: int | None = None, autos: bool = False,
                     out: np.ndarray | None = None):
        """fx_process_host: HOST uint8 buffers in, HOST complex64 rows out (H2D/D2H inside)."""
        if n_blocks is None:
            n_blocks = raw0.size // (2 * self.num_samp)
        for r in (raw0, raw1):
            if r.dtype != np.uint8 or not r.flags.c_contiguous or r.size < 2 * self.num_samp * n_blocks:
                raise ValueError("raw IQ must be contiguous uint8 of at least n_blocks*2*num_samp bytes")
        x = np.empty((n_blocks, self.nbins), dtype=np.complex64) if out is None else out
        a0 = np.empty((n_blocks, self.nbins), dtype=np.float32) if autos else None
        a1 = np.empty((n_blocks, self.nbins), dtype=np.float32) if autos else None
        rc = self.lib.fx_process_host(self.h, raw0.ctypes.data, raw1.ctypes.data, n_blocks, x.ctypes.data,
                                      a0.ctypes.data if autos else None, a1.ctypes.data if autos else None)
        self._check(rc, "fx_process_host")
        return (x, a0, a1) if autos else x

    def copy_probe(self, raw0: np.ndarray, raw1: np.ndarray, n_blocks: int, out: np.ndarray):
        """fx_copy_probe: the H2D/D2H traffic of process_host with no kernels (the roof of that path)."""
        self._check(self.lib.fx_copy_probe(self.h, raw0.ctypes.data, raw1.ctypes.data, n_blocks, out.ctypes.data),
                    "fx_copy_probe")

    # ---- pieces ---------------------------------------------------------------
    def pfb(self, x) -> torch.Tensor:
        """_spectrometer_poly: complex input of num_samp samples -> (P, N) complex64."""
        if isinstance(x, torch.Tensor) and x.dtype == torch.uint8:
            p = self._raw(x, 1)
            out = torch.empty((self.frames_per_block, self.nbins), dtype=torch.complex64, device=self.tdev)
            self._enter()
            self._check(self.lib.fx_pfb_u8(self.h, p, out.data_ptr()), "fx_pfb_u8")
            self._exit()
            return out
        xt = self._as_c64(x)
        if xt.numel() != self.num_samp:
            raise ValueError("input length must equal the engine's num_samp")
        out = torch.empty((self.frames_per_block, self.nbins), dtype=torch.complex64, device=self.tdev)
        self._enter()
        self._check(self.lib.fx_pfb_c64(self.h, xt.data_ptr(), out.data_ptr()), "fx_pfb_c64")
        self._exit()
        return out

    def _as_c64(self, x) -> torch.Tensor:
        if isinstance(x, torch.Tensor):
            return x.to(device=self.tdev, dtype=torch.complex64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(np.asarray(x).astype(np.complex64))).to(self.tdev)

    def lag(self, a, b, n_blocks: int = 1):
        """Lag search of effex.py:605-622.  Returns (n, imax, xprev, xbest, xnext)."""
        imax = C.c_int64()
        nb = (C.c_float * 3)()
        if isinstance(a, torch.Tensor) and a.dtype == torch.uint8:
            p0, p1 = self._raw(a, n_blocks), self._raw(b, n_blocks)
            self._enter()
            rc = self.lib.fx_lag_u8(self.h, p0, p1, n_blocks, C.byref(imax), nb)
        else:
            if len(a) != len(b):
                raise AssertionError('Algorithm assumes input complex timeseries are of equal length.')
            ta, tb = self._as_c64(a), self._as_c64(b)
            if ta.numel() != self.num_samp * n_blocks:
                raise ValueError("input length must equal n_blocks * num_samp")
            self._enter()
            rc = self.lib.fx_lag_c64(self.h, ta.data_ptr(), tb.data_ptr(), n_blocks, C.byref(imax), nb)
        self._check(rc, "fx_lag")
        return self.num_samp, int(imax.value), float(nb[0]), float(nb[1]), float(nb[2])

    def lag_fft_len(self) -> int:
        return int(self.lib.fx_lag_fft_len(self.h))

    def lag_accumulate(self, a, b, n_blocks: int = 1, xacc: torch.Tensor | None = None, first: bool = True):
        """fx_lag_accumulate_*: xacc[M] (complex64) (=|+=) sum over the block pairs of FFT(a)*conj(FFT(b))."""
        if xacc is None:
            xacc = torch.empty(self.lag_fft_len(), dtype=torch.complex64, device=self.tdev)
            first = True
        if isinstance(a, torch.Tensor) and a.dtype == torch.uint8:
            p0, p1 = self._raw(a, n_blocks), self._raw(b, n_blocks)
            self._enter()
            rc = self.lib.fx_lag_accumulate_u8(self.h, p0, p1, n_blocks, xacc.data_ptr(), 1 if first else 0)
        else:
            ta, tb = self._as_c64(a), self._as_c64(b)
            if ta.numel() != self.num_samp * n_blocks or tb.numel() != ta.numel():
                raise ValueError("input length must equal n_blocks * num_samp")
            self._enter()
            rc = self.lib.fx_lag_accumulate_c64(self.h, ta.data_ptr(), tb.data_ptr(), n_blocks, xacc.data_ptr(),
                                                1 if first else 0)
        self._check(rc, "fx_lag_accumulate")
        self._exit()
        return xacc

    def lag_finish(self, xacc: torch.Tensor):
        """fx_lag_finish: (n, imax, xprev, xbest, xnext) from an accumulated cross-spectrum."""
        imax = C.c_int64()
        nb = (C.c_float * 3)()
        self._enter()
        self._check(self.lib.fx_lag_finish(self.h, xacc.data_ptr(), C.byref(imax), nb), "fx_lag_finish")
        return self.num_samp, int(imax.value), float(nb[0]), float(nb[1]), float(nb[2])

    # ---- measurement -----------------------------------------------------------
    def reset_counters(self):
        self._check(self.lib.fx_reset_counters(self.h), "fx_reset_counters")

    def kernel_launches(self) -> int:
        return int(self.lib.fx_kernel_launches(self.h))

    def enable_timing(self, on: bool = True):
        self._check(self.lib.fx_enable_timing(self.h, 1 if on else 0), "fx_enable_timing")

    def dominant_kernel_time(self):
        ms = C.c_double()
        n = C.c_int64()
        self._check(self.lib.fx_dominant_kernel_time(self.h, C.byref(ms), C.byref(n)), "fx_dominant_kernel_time")
        return float(ms.value), int(n.value)
