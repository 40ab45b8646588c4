"""Per-phase clock64 timeline of the tensor-core recurrent kernels (CTA 0), layer-1-like shape."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ss_asr_b200 import _lib, functional as Fk
from ss_asr_b200.asr import pBLSTM
lib = _lib.load()
dev = 'cuda'
B, T, K, S = 256, 512, 1024, 256
torch.manual_seed(0)
m = pBLSTM(K, S).to(dev)
m.precision = 'bf16'
x = torch.randn(B, T, K, device=dev, requires_grad=True)
lens = [T] * B
dbg = torch.zeros(T, 12, dtype=torch.int64, device=dev)
for which in ('fwd', 'bwd'):
    out, _, _ = m(x, state_len=lens, pack_input=True)
    torch.cuda.synchronize()
    if which == 'fwd':
        lib.ssasr_rec_tc_set_debug(dbg.data_ptr())
        out, _, _ = m(x, state_len=lens, pack_input=True)
        torch.cuda.synchronize()
        lib.ssasr_rec_tc_set_debug(None)
    else:
        lib.ssasr_rec_tc_set_debug(dbg.data_ptr())
        out.sum().backward()
        torch.cuda.synchronize()
        lib.ssasr_rec_tc_set_debug(None)
    d = dbg.cpu()
    print(which, 'stamps (cycles rel. to e0 of each step), steps 100..104:')
    for s in range(100, 105):
        r = d[s]
        print('  step', s, [int(v - r[0]) for v in r], ' next e0 - e0 =', int(d[s + 1][0] - r[0]))
    per = (d[400][0] - d[100][0]).item() / 300
    print(which, 'avg cycles/step', per)
    dbg.zero_()
