#!/usr/bin/env python
"""Generate golden fixtures by EXECUTING THE REFERENCE'S OWN SOURCE FILE.

Run in the build container only (`/root/reference` does not exist on the GPU
box):   python tests/golden/make_golden.py

`effex/effex.py` imports cupy, cusignal, rtlsdr and (through post_process)
matplotlib, none of which are installed here and all of which are unpinned
third-party code.  This script registers numpy-backed stand-ins for those
modules, imports the UNMODIFIED `/root/reference/effex/effex.py`, and calls
its methods (`_spectrometer_poly`, `_pfb_xcorr`, `_estimate_delay_gaussian`,
`_estimate_delay`, `_write_metadata`, the DC-removal expression is restated in
`ref_dc`) on seeded inputs.  What the stand-ins provide:

  cupy      -> numpy (array/zeros/fft/exp/conj/pi/argmax/abs/asnumpy/...)
  cusignal  -> scipy.signal.get_window / firwin (cuSignal's are ports of
               these); get_shared_mem -> np.zeros; filtering.channelize_poly
               -> tests/golden/cusignal_standin.py, a thread-by-thread loop
               form of cuSignal's channelizer kernel that shares NO code with
               the oracle (so the spec0 / xspec fixtures pin the oracle's
               channelizer -- tap order, zero fill, conjugations -- from outside)
  rtlsdr    -> a dummy RtlSdr that accepts attribute writes (no USB)
  matplotlib-> empty module (only imported, never used here)

Outputs (small, committed): tests/golden/ref_*.npz, tests/golden/ref_meta_*.csv
"""
import os
import sys
import types
import logging

import numpy as np
import scipy.signal

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import fx_oracle as orc  # noqa: E402

REF = "/root/reference/effex"


def install_shims():
    cp = types.ModuleType("cupy")
    for name in ("array", "zeros", "exp", "conj", "pi", "argmax", "abs", "real", "linspace",
                 "cos", "sin", "roll", "complex128", "float64", "asarray"):
        setattr(cp, name, getattr(np, name))
    cp.fft = np.fft
    cp.random = np.random
    cp.asnumpy = np.asarray
    sys.modules["cupy"] = cp

    cs = types.ModuleType("cusignal")
    cs.get_window = scipy.signal.get_window
    cs.firwin = scipy.signal.firwin
    cs.get_shared_mem = lambda n, dtype=np.complex128: np.zeros(n, dtype=dtype)
    filt = types.ModuleType("cusignal.filtering")
    sys.path.insert(0, HERE)
    import cusignal_standin
    filt.channelize_poly = cusignal_standin.channelize_poly      # NOT the oracle's
    cs.filtering = filt
    sys.modules["cusignal"] = cs
    sys.modules["cusignal.filtering"] = filt

    rt = types.ModuleType("rtlsdr")

    class RtlSdr:
        def __init__(self, *a, **k):
            pass

        def close(self):
            pass
    rt.RtlSdr = RtlSdr
    sys.modules["rtlsdr"] = rt

    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


def make_correlator(fx, **kw):
    logging.disable(logging.CRITICAL)
    cwd = os.getcwd()
    os.chdir("/tmp")            # the ctor opens log_effex.log in the CWD
    try:
        return fx.Correlator(**kw)
    finally:
        os.chdir(cwd)


def ref_dc(x):
    # effex/effex.py:394 (the expression, applied to one channel)
    return (x.real - x.real.mean()) + 1j * (x.imag - x.imag.mean())


def noise(rng, n, scale):
    return rng.normal(size=n, scale=scale) + 1j * rng.normal(size=n, scale=scale)


def main():
    install_shims()
    sys.path.insert(0, REF)
    import effex as fx  # the reference, unmodified

    rng = np.random.default_rng(77777)      # seed from tests/test_effex.py:10

    # ---- case A: small spectrum-mode block pair, with a DC offset and a delay
    for tag, S, N, bw, fc, tau_samples in (("a", 4096, 256, 2.4e6, 1.4204e9, 5),
                                           ("b", 8192, 1024, 3.2e6, 1.0e8, 0)):
        cor = make_correlator(fx, run_time=1, bandwidth=bw, frequency=fc,
                              num_samp=S, nbins=N, mode="spectrum")
        g = noise(rng, S + 64, 0.25)
        x0 = g[tau_samples:tau_samples + S] + noise(rng, S, 0.1) + (0.013 - 0.021j)
        x1 = g[:S] + noise(rng, S, 0.1) + (-0.008 + 0.004j)
        # quantise like an RTL-SDR so the same bytes can be fed to the GPU path
        def q(x):
            b = np.empty(2 * len(x), dtype=np.uint8)
            b[0::2] = np.clip(np.rint(127.5 + 127.5 * x.real), 0, 255)
            b[1::2] = np.clip(np.rint(127.5 + 127.5 * x.imag), 0, 255)
            return b
        raw0, raw1 = q(x0), q(x1)
        u0, u1 = orc.unpack_iq(raw0), orc.unpack_iq(raw1)
        cor.gpu_iq_0 = ref_dc(u0)
        cor.gpu_iq_1 = ref_dc(u1)
        out = {"raw0": raw0, "raw1": raw1, "S": S, "N": N, "bw": bw, "fc": fc,
               "window": np.asarray(cor.window)}
        out["spec0"] = cor._spectrometer_poly(np.array(cor.gpu_iq_0), cor.ntaps, cor.nbins, cor.window)
        cor.calibrated_delay = 0
        out["xspec_tau0"] = np.asarray(cor._pfb_xcorr())
        cor._calibrate_task()
        out["calibrated_delay"] = cor.calibrated_delay
        out["xspec_cal"] = np.asarray(cor._pfb_xcorr())
        out["delay_gauss"] = cor._estimate_delay_gaussian(cor.gpu_iq_0, cor.gpu_iq_1, bw)
        cor.mode = "continuum"
        out["vis_continuum"] = np.asarray(cor._pfb_xcorr())
        cor.mode = "test"
        out["delay_test_mode"] = cor._estimate_delay(cor.gpu_iq_0, cor.gpu_iq_1, bw)
        out["test_delay_offset"] = cor.test_delay_offset
        out["test_delay_sweep_step"] = cor.test_delay_sweep_step
        cor.mode = "spectrum"
        # CSV header written by the reference's own _write_metadata
        cor.output_file = os.path.join(HERE, f"ref_meta_{tag}.csv")
        cor._write_metadata()
        with open(cor.output_file, "a") as fh:      # effex/effex.py:689,693
            np.savetxt(fh, [np.asarray(out["xspec_cal"])], delimiter=',')
        np.savez_compressed(os.path.join(HERE, f"ref_case_{tag}.npz"), **out)
        print(tag, "delay", out["calibrated_delay"] * bw, "samples; |X| max", np.abs(out["xspec_cal"]).max())

    # ---- case C: continuum-mode metadata (no frequency row)
    cor = make_correlator(fx, run_time=2, bandwidth=2.4e6, frequency=1.4204e9,
                          num_samp=4096, nbins=256, mode="continuum")
    cor.output_file = os.path.join(HERE, "ref_meta_c.csv")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cor._write_metadata()

    # ---- case D: num_samp clamp / validation behaviour of the ctor
    clamp = {}
    for v in (100, 256, 5000, 2 ** 18, 2 ** 20):
        c = make_correlator(fx, num_samp=v, nbins=16)
        clamp[str(v)] = c.num_samp
    np.savez(os.path.join(HERE, "ref_clamp.npz"), **clamp)
    print("clamp", clamp)


if __name__ == "__main__":
    main()
