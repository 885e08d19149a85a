#!/usr/bin/env python
"""cuFFT as the BAR for the FFT stage ("beat cuFFT C2C batched 4096 on the same box").
The reference runs cuFFT Z2Z 64 x 4096 per channel per block (effex.py:553) and three 2^19-point transforms per
calibration (:611-613).  Timed here with CUDA events on the same GPU, complex64 (the float32 the product
computes in; Z2Z would be ~32x slower on this part's FP64 rate):

  A. cuFFT C2C batch 70 400 x 4096 (both channels of 550 blocks x 64 frames)  vs  bigfft::tail_kernel over
     the same 35 200 channel-PACKED frames resident in HBM (stages A-C of the fused kernel + the X-engine, fed
     from Z: the FFT + X part of the fused kernel and nothing else);
  B. cuFFT C2C 2 x 2^19 per block pair, 92 block pairs (+ the multiply-accumulate and the inverse in torch)
     vs  fx_lag_accumulate_u8 (unpack + zero-pad + transforms + accumulate) and fx_lag_u8.
Writes a summary to stdout (tee it into gpurun_out/ and copy to profiles/)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from effex_b200 import synth
from effex_b200.engine import FxEngine


def ev_time(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(torch.cuda.get_device_name(0), "torch", torch.__version__)
# ---- A: batched 4096-point transforms -------------------------------------------------------------------
frames = 550 * 64
x = torch.randn(2 * frames, 4096, dtype=torch.complex64, device="cuda")
y = torch.empty_like(x)
t_cufft = ev_time(lambda: torch.fft.fft(x, dim=1, out=y), 10)
gb = 2 * x.numel() * 8 / 1e9
print(f"A. cuFFT C2C 70400 x 4096 (complex64, out of place): {t_cufft*1e3:8.1f} us   ({gb / (t_cufft*1e-3):6.0f} GB/s of its own traffic: {gb:.2f} GB)")
del x, y
S, N, nb = 262144, 8192, 550           # 8192 bins: the tail kernel sees nb * 32 frames * 2 virtual blocks = 35200 transforms of 4096
raw0, raw1 = synth.tiled_recording(nb, S, base_blocks=4)
d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
eng = FxEngine(S, N, 4, max_blocks=nb)
out = (torch.empty((nb, N), dtype=torch.complex64, device="cuda"), None, None)
for _ in range(3):
    eng.process(d0, d1, nb, out=out)
eng.sync(); eng.reset_counters(); eng.enable_timing(True)
for _ in range(10):
    eng.process(d0, d1, nb, out=out)
eng.sync()
ms, n = eng.dominant_kernel_time()
t_tail = ms / 10                      # per pass: the pass walks Z in chunks of <= 1 GiB, one tail launch per chunk
print(f"   bigfft::tail_kernel, 35200 packed frames from Z (2.31 GB read, {n // 10} launches): {t_tail*1e3:8.1f} us  -> {t_cufft / t_tail:5.2f}x cuFFT"
      f"  (FFT of both channels + X-engine; cuFFT's time is the transforms alone)")
eng.close()
del d0, d1, out
# ---- B: the 2^19-point lag transforms ----------------------------------------------------------------------
n, nb = 262144, 92
M = 2 * n
raw0, raw1 = synth.tiled_recording(nb, n, base_blocks=4, delay=37, seed=99)
d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
a = torch.randn(nb, n, dtype=torch.complex64, device="cuda")
b = torch.randn(nb, n, dtype=torch.complex64, device="cuda")
t_fwd = ev_time(lambda: (torch.fft.fft(a, n=M, dim=1), torch.fft.fft(b, n=M, dim=1)), 5)
def torch_lag():
    A, B = torch.fft.fft(a, n=M, dim=1), torch.fft.fft(b, n=M, dim=1)
    xc = torch.fft.ifft((A * B.conj()).sum(dim=0))
    return torch.argmax(xc.abs())
t_all = ev_time(torch_lag, 5)
eng = FxEngine(n, 4096, 1, max_blocks=nb)
xacc = eng.lag_accumulate(d0, d1, nb)
t_acc = ev_time(lambda: eng.lag_accumulate(d0, d1, nb, xacc=xacc, first=True), 10)
eng.lag(d0, d1, nb)
t0 = time.perf_counter()
for _ in range(10):
    r = eng.lag(d0, d1, nb)
t_lag = (time.perf_counter() - t0) / 10 * 1e3
print(f"B. cuFFT C2C 184 x 2^19 (zero-padded from complex64 input): {t_fwd*1e3:8.1f} us; with multiply-accumulate, inverse, argmax in torch: {t_all*1e3:8.1f} us")
print(f"   fx_lag_accumulate_u8, 92 block pairs (unpack + DC + pad + transforms + accumulate): {t_acc*1e3:8.1f} us -> {t_fwd / t_acc:5.2f}x the cuFFT forward transforms alone")
print(f"   fx_lag_u8, 92 block pairs, synchronous with the result on the host: {t_lag*1e3:8.1f} us (lag {r[0]-r[1]}) -> {t_all / t_lag:5.2f}x the torch/cuFFT chain")
a1, b1 = a[:1].contiguous(), b[:1].contiguous()
t1 = ev_time(lambda: (torch.fft.fft(a1, n=M, dim=1), torch.fft.fft(b1, n=M, dim=1)), 20)
t1a = ev_time(lambda: eng.lag_accumulate(d0, d1, 1, xacc=xacc, first=True), 20)
print(f"   one block pair: cuFFT 2 x 2^19: {t1*1e3:8.1f} us ; fx_lag_accumulate_u8: {t1a*1e3:8.1f} us")
eng.close()
