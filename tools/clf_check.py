"""Consumer classifier on the GPU box: fused bf16 form vs the fp32 module, and its time per batch of 64."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import synthetic as S  # noqa: E402

torch.backends.cudnn.benchmark = True
clf = ap.CifarResNeXt(nlabels=10, in_channels=1)
clf.load_state_dict(S.resnext_state_dict(4321))
clf = clf.cuda().eval()
fused = ap.FusedResNeXt(clf).cuda()
x = torch.randn(64, 1, 32, 32, device="cuda") * 20 - 30
with torch.no_grad():
    a, b = clf(x), fused(x)
print("rel", float((a - b).norm() / a.norm()), "argmax agree", float((a.argmax(1) == b.argmax(1)).float().mean()))
for _ in range(5):
    fused(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    fused(x)
e1.record()
torch.cuda.synchronize()
print("fused ms per batch of 64:", e0.elapsed_time(e1) / 20)
