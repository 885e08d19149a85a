#!/usr/bin/env python
"""Step time of the canonical workload (550 blocks, 4096 bins) in the three flavours a rank can run:
rows only (fx_process), rows + local accumulators (fx_process_acc), rows + reduce (fx_process_reduce,
world of one: same kernels as a multi-GPU rank).  Run under gpurun; with --ncu it runs few steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from effex_b200 import synth
from effex_b200.engine import FxEngine

S, N, NB = 262144, 4096, 550
steps = 3 if "--ncu" in sys.argv else 50
raw0, raw1 = synth.tiled_recording(NB, S, base_blocks=4)
d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
eng = eng0 = FxEngine(S, N, 4, max_blocks=NB)
tok = eng.comm_export(1)
eng.comm_attach(0, 1, [tok])
out = (torch.empty((NB, N), dtype=torch.complex64, device="cuda"), None, None)
acc = eng.new_accumulators()
modes = {
    "rows": lambda: eng.process(d0, d1, NB, out=out, inputs_ready=True),
    "rows+acc": lambda: eng.process(d0, d1, NB, out=out, acc=acc, inputs_ready=True),
    "rows+reduce": lambda: eng.process_reduce(d0, d1, NB, out=out, acc=acc, root=0, inputs_ready=True),
}
lean = FxEngine(S, N, 4, max_blocks=NB, cross_only=True)
lean.comm_attach(0, 1, [lean.comm_export(1)])
acc2 = lean.new_accumulators()
modes["rows+acc (cross only)"] = lambda: lean.process(d0, d1, NB, out=out, acc=acc2, inputs_ready=True)
modes["rows+reduce (cross only)"] = lambda: lean.process_reduce(d0, d1, NB, out=out, acc=acc2, root=0, inputs_ready=True)
for name, fn in modes.items():
    eng = lean if "cross only" in name else eng0
    for _ in range(3):
        fn()
    eng.sync(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.reset_counters(); eng.enable_timing(True)
    e0.record()
    for _ in range(steps):
        fn()
    eng.comm_fence()
    torch.cuda.current_stream().wait_stream(eng.stream)
    e1.record()
    eng.sync(); torch.cuda.synchronize()
    ms, n = eng.dominant_kernel_time()
    eng.enable_timing(False)
    print(f"{name:26s} step {e0.elapsed_time(e1) / steps * 1e3:8.1f} us   fused kernel {ms / n * 1e3:8.1f} us   launches/step {eng.kernel_launches() / steps:.1f}")
eng0.close(); lean.close()
