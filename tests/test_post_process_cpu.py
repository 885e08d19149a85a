import numpy as np

from effex_b200 import post_process as pp


def test_panels_and_envelope_fit():
    v = np.array([1 + 1j, -2j, 3.0])
    amp, phase, re, im = pp.panels(v)
    np.testing.assert_allclose(amp, np.abs(v))
    np.testing.assert_allclose(phase, np.angle(v))
    # the TEST-mode fit recovers a synthetic sinc^2 envelope (no plotting)
    step, bw = 3.5e-10, 2.4e6
    n = 400
    tau = np.arange(-n // 2, n // 2) * step
    truth = (2.0e-3, 5.0e-8, 2.2e6, 0.0)
    vis = pp.envelope(tau, *truth).astype(complex)      # the reference fits the squared model to |V| itself
    fit = pp.fit_interferometer_model(vis, step, bw, 1.4204e9, show=False)
    assert abs(abs(fit[0]) - truth[0]) / truth[0] < 1e-3
    pp.post_process(vis, bw, 1.4204e9, 1024, 'test', omit_plot=True)        # no matplotlib needed
