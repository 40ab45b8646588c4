"""Two C4 training steps (B=256, T=512, bf16 path, deferred weight gradients) for ncu captures of the recurrent / GEMM kernels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as BN
from ss_asr_b200 import functional as Fk
from ss_asr_b200.functional import asr_loss
from ss_asr_b200.optim import FusedAdadelta
dev = torch.device('cuda', 0)
m = BN.fresh_model(dev); m.train_precision = 'bf16'; m.train(); m.att_on_device = True
opt = FusedAdadelta(m.parameters(), lr=1.0, eps=1e-8)
Fk.set_overlap_wgrad(True)
x, lens, y = BN.synth_batch(256, 512, 80, 40)
xd, yd = x.to(dev), y.to(dev)
ans = int(max((y != 0).sum(-1) + 1)) - 1
for _ in range(2):
    opt.zero_grad(set_to_none=True)
    _, logits, _ = m(xd, ans, teacher=yd, state_len=lens)
    asr_loss(logits, yd).backward()
    opt.step_clipped(5.0)
torch.cuda.synchronize()
print('ok')
