"""A/B of the greedy loop schedule (SSASR_DECODE_DUAL) on NU utterances of the C3 recipe.  GPU only."""
import os, sys, torch
sys.path.insert(0, '/root/repo')
from ss_asr_b200.asr import ASR
dev='cuda'
torch.manual_seed(1)
m = ASR(50, 256, 256, 128, 80, 0.9).to(dev).eval()
g = torch.Generator().manual_seed(4321)
NU = int(os.environ.get('NU', '1000'))
Ts = sorted([int(v) for v in torch.randint(256, 513, (NU,), generator=g)], reverse=True)
xb = torch.zeros(len(Ts), Ts[0], 80)
for i, t in enumerate(Ts):
    xb[i, :t] = torch.randn(t, 80, generator=g)
xb = xb.to(dev)
ref=None
for env in ('0', '1', '0', '1'):
    os.environ['SSASR_DECODE_DUAL'] = env
    m.decode_batch(xb, Ts, precision='tf32x3')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ids = m.decode_batch(xb, Ts, precision='tf32x3'); e1.record(); torch.cuda.synchronize()
    ref = ref or ids
    print('dual', env, '%.2f ms' % e0.elapsed_time(e1), ids == ref, flush=True)
