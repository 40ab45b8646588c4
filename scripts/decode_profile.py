"""C3 greedy decode: per-family time of one decode_batch call for the exact paths and several Listener group sizes."""
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
ref = None
for prec, chunk in (('fp32', 0), ('tf32x3', 0), ('tf32x3', 512), ('tf32x3', 256), ('tf32x3', 128), ('fp32', 256)):
    m.decode_encoder_chunk = chunk
    m.decode_batch(xb, Ts, precision=prec)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ids = m.decode_batch(xb, Ts, precision=prec); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ref = ref or ids
    lib.ssasr_profile_enable(1); _lib.profile_read()
    m.decode_batch(xb, Ts, precision=prec)
    p = _lib.profile_read(); lib.ssasr_profile_enable(0)
    print(prec, 'chunk', chunk, '%.1f ms' % ms, 'identical' if ids == ref else 'DIFFERENT',
          {k: (round(v[0], 2), v[1]) for k, v in sorted(p.items(), key=lambda kv: -kv[1][0]) if v[1]})
