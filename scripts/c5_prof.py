"""One C5 training step (T=1600, S_enc=512, B=32) after a warm-up step (for ncu captures of rec_wide.cu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as BN
from ss_asr_b200.functional import asr_loss
dev = torch.device('cuda', 0)
m = BN.fresh_model(dev, S_enc=512); m.train_precision = 'bf16'; m.train(); m.att_on_device = True
x, lens, y = BN.synth_batch(32, 1600, 80, 100)
xd, yd = x.to(dev), y.to(dev)
ans = int(max((y != 0).sum(-1) + 1)) - 1
for _ in range(2):
    m.zero_grad(set_to_none=True)
    _, logits, _ = m(xd, ans, teacher=yd, state_len=lens)
    asr_loss(logits, yd).backward()
torch.cuda.synchronize()
print('ok')
