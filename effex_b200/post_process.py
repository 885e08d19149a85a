"""Post-processing of a recorded run: the reference's `post_process.py` (plots of amplitude, phase,
real and imaginary parts; a sinc-envelope fit in TEST mode).  Pure host code with no GPU content
(SURVEY 8f rank 4); matplotlib is imported lazily because it is an optional dependency.

    python -m effex_b200.post_process visibilities_YYYYmmdd-HHMMSS.csv
"""
from __future__ import annotations

import argparse

import numpy as np


def panels(visibilities):
    """The four quantities the reference plots (post_process.py:13-16)."""
    v = np.asarray(visibilities)
    return np.sqrt(np.real(v * np.conj(v))), np.angle(v), np.real(v), np.imag(v)


def envelope(tau, amp, tau0, dnu, slope):
    """Delay-domain bandwidth pattern used by the reference's fit (post_process.py:115-123,
    Thompson, Moran & Swenson eq. 2.4 cast in delay): (amp*sinc(pi*(tau+tau0)*dnu) + slope*tau)^2."""
    return (amp * np.sinc(np.pi * (tau + tau0) * dnu) + slope * tau) ** 2


def fit_interferometer_model(raw_output, delay_step, bandwidth, center_freq, show=True):
    from scipy.optimize import curve_fit
    amp, _, _, _ = panels(raw_output)
    samples = np.arange(-len(amp) // 2, len(amp) // 2)
    delay = samples * delay_step
    p0 = [np.max(amp) ** .5, 5.84e-8, bandwidth, 0]
    pfit, _ = curve_fit(envelope, delay, amp, p0)
    print(pfit)
    if show:
        import matplotlib.pyplot as plt
        _, ax = plt.subplots()
        ax.plot(delay, amp, label='measurement')
        ax.plot(delay, envelope(delay, *pfit), label='sinc envelope fit')
        ax.set_xlabel('Delay (s)')
        ax.set_ylabel('Amplitude (adu)')
        ax.legend()
        plt.show()
    return pfit


def visualize(visibilities, rate, fc, nfft, mode, test_delay_sweep_step=0):
    import matplotlib.pyplot as plt
    amp, phase, re, im = panels(visibilities)
    scalar = mode in ('continuum', 'test')
    fig, axes = plt.subplots(nrows=2, ncols=2, sharex='all', sharey='none' if scalar else 'all')
    if scalar:
        x = np.arange(len(amp))
        xlabel = 'Sample #'
        if test_delay_sweep_step:
            x, xlabel = x * test_delay_sweep_step * 1e9, 'Delay (ns)'
        for ax, y, title, ylabel in ((axes[0][0], amp, 'Complex Cross-Correlation Amplitude', 'Amplitude (uncalibrated)'),
                                     (axes[1][0], phase, 'Complex Cross-Correlation Phase', 'Phase'),
                                     (axes[1][1], im, 'Complex Cross-Correlation Imag', 'Amplitude')):
            ax.plot(x, y)
            ax.set_xlabel(xlabel); ax.set_ylabel(ylabel); ax.set_title(title)
        axes[0][1].plot(x, re, label='real part')
        axes[0][1].plot(x, im, alpha=0.5, label='imag_part')
        axes[0][1].set_xlabel(xlabel); axes[0][1].set_ylabel('Amplitude')
        axes[0][1].set_title('Complex Cross-Correlation Real & Imag'); axes[0][1].legend(loc='best')
    else:
        freqs = np.fft.fftshift(np.fft.fftfreq(nfft, d=1 / rate)) + fc
        rows = np.arange(np.atleast_2d(visibilities).shape[0])
        stride = max(1, int(rows.max()) // 50) if rows.size and rows.max() > 50 else 1   # <= ~50 rows on screen
        X, Y = np.meshgrid(freqs, rows[::stride])
        for ax, z, title in ((axes[0][0], amp, 'Complex Cross-Correlation Amplitude'), (axes[0][1], re, 'Real part of XCorrs'),
                             (axes[1][0], phase, 'Complex Cross-Correlation Phase'), (axes[1][1], im, 'Imag part of XCorrs')):
            m = ax.pcolormesh(X, Y, np.atleast_2d(z)[::stride, :], shading='auto', cmap='viridis')
            if z is phase:
                m.set_clim(-np.pi, np.pi)
            ax.set_xlabel('Frequency (Hz)'); ax.set_ylabel('Sample #'); ax.set_title(title)
            fig.colorbar(m, ax=ax)
    fig.tight_layout()
    plt.show()


def post_process(raw_output, rate, fc, nfft, mode, omit_plot, test_delay_sweep_step=0):
    """Same signature as the reference's post_process (post_process.py:150)."""
    if not omit_plot:
        visualize(raw_output, rate, fc, nfft, mode, test_delay_sweep_step=test_delay_sweep_step)
        if mode == 'test':
            fit_interferometer_model(raw_output, test_delay_sweep_step, rate, fc)


def main(argv=None):
    parser = argparse.ArgumentParser(description='Re-plot an effex .csv file.',
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('filename', type=str, help='(str) output visibilities.csv file from effex.')
    args = parser.parse_args(argv)
    from .csvio import read_rows
    meta, rows = read_rows(args.filename)
    mode = meta['mode'].lower()
    # the reference's re-plotter assumes a sweep step of (1/f)/10 although the correlator steps by
    # (1/f)/2; the step actually used by the run is the latter
    step = (1 / float(meta['frequency'])) / 2. if mode == 'test' else 0
    if mode != 'spectrum':
        rows = rows.reshape(-1)
    post_process(rows, float(meta['bandwidth']), float(meta['frequency']), int(meta['resolution']), mode, False,
                 test_delay_sweep_step=step)


if __name__ == '__main__':
    main()
