"""One C3 greedy decode (1000 utterances, exact split-operand path) for ncu: launch list and full captures of its kernels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ss_asr_b200.asr import ASR
dev = 'cuda'
torch.manual_seed(1)
m = ASR(50, 256, 256, 128, 80, 0.9).to(dev).eval()
g = torch.Generator().manual_seed(4321)
Ts = sorted([int(v) for v in torch.randint(256, 513, (1000,), generator=g)], reverse=True)
xb = torch.zeros(len(Ts), Ts[0], 80)
for i, t in enumerate(Ts):
    xb[i, :t] = torch.randn(t, 80, generator=g)
xb = xb.to(dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    ids = m.decode_batch(xb, Ts, precision='tf32x3')
torch.cuda.synchronize()
print('decoded', len(ids), 'utterances,', m.last_decode_steps, 'steps')
