import sys; sys.path.insert(0, "/root/repo")
import torch
from effex_b200 import synth
from effex_b200.engine import FxEngine
S, N, nb = 2**24, 65536, 2
raw0, raw1 = synth.tiled_recording(nb, S, base_blocks=2)
d0, d1 = torch.from_numpy(raw0).cuda(), torch.from_numpy(raw1).cuda()
eng = FxEngine(S, N, 4, max_blocks=nb)
out = (torch.empty((nb, N), dtype=torch.complex64, device="cuda"), None, None)
for _ in range(2): eng.process(d0, d1, nb, out=out)
eng.sync()
