"""Two forward + backward calls of the attend-and-spell loop at the C4 decoder shapes (for ncu captures of spell_cl.cu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ss_asr_b200.asr import ASR
dev = 'cuda'
torch.manual_seed(1)
m = ASR(50, 256, 256, 128, 80, 0.9).to(dev).train()
B, Tp, U = 256, 64, 41
g = torch.Generator().manual_seed(1)
enc = (0.3 * torch.randn(B, Tp, 512, generator=g)).to(dev)
lens = sorted([int(v) for v in torch.randint(48, 65, (B,), generator=g)], reverse=True)
tok = torch.randint(3, 50, (B, U), generator=g).to(torch.int32).to(dev)
for _ in range(2):
    e = enc.clone().requires_grad_(True)
    logits, att, toks = m._spell(e, lens, tok.clone(), [0] * U, 'bf16')
    logits.backward(torch.ones_like(logits) * 1e-3)
torch.cuda.synchronize()
print('ok')
