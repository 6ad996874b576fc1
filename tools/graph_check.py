"""CUDA-graph capture of the purification call (small batches are launch-latency bound): capture
Engine.ddpm_purify on a side stream, replay, compare with eager, time both.  GPU box only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import synthetic as S  # noqa: E402

m = ap.WaveNet_Speech_Commands(**S.DEFAULT_WAVENET_CONFIG)
m.load_state_dict(S.diffwave_state_dict(1234))
m = m.cuda().eval()
eng = m.engine()
for B in (1, 2, 4, 8):
    x = S.waveforms(B, 16000, seed=B).cuda()
    z = S.noise((2, B, 1, 16000), seed=3).cuda()
    eager = eng.ddpm_purify(x, 2, z=z)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        eng.ddpm_purify(x, 2, z=z)  # warm-up on the capture stream
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g, stream=s):
        out = eng.ddpm_purify(x, 2, z=z)
    g.replay()
    torch.cuda.synchronize()
    same = torch.equal(out, eager)

    def timeit(fn, n=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    te = timeit(lambda: eng.ddpm_purify(x, 2, z=z))
    tg = timeit(g.replay)
    print("B=%d graph==eager: %s   eager %.3f ms   graph %.3f ms" % (B, same, te, tg))
