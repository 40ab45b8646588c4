import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ss_asr_b200 import _lib
from ss_asr_b200.asr import ASR
lib = _lib.load()
dev = 'cuda'
torch.manual_seed(1)
m = ASR(50, 256, 256, 128, 80, 0.9).to(dev).eval()
g = torch.Generator().manual_seed(4321)
Ts = sorted([int(v) for v in torch.randint(256, 513, (1000,), generator=g)], reverse=True)
xb = torch.zeros(len(Ts), Ts[0], 80)
for i, t in enumerate(Ts):
    xb[i, :t] = torch.randn(t, 80, generator=g)
xb = xb.to(dev)
for prec in ('fp32', 'tf32x3'):
    m.decode_batch(xb, Ts, precision=prec)
    lib.ssasr_profile_enable(1); _lib.profile_read()
    m.decode_batch(xb, Ts, precision=prec)
    p = _lib.profile_read(); lib.ssasr_profile_enable(0)
    print(prec, {k: (round(v[0], 2), v[1]) for k, v in sorted(p.items(), key=lambda kv: -kv[1][0]) if v[1]})
