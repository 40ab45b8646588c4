"""CUDA-event timing of one pBLSTM layer (forward and backward calls) at the C4 layer shapes; no debug stamps."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ss_asr_b200 import _lib
from ss_asr_b200.asr import pBLSTM
lib = _lib.load()
dev = 'cuda'
tag = 'cluster=%s flags=%s' % (os.environ.get('SSASR_REC_CLUSTER', '1'), os.environ.get('SSASR_CL_FLAGS', '0'))
for (B, T, K, S) in [(256, 512, 80, 256), (256, 256, 1024, 256), (256, 128, 1024, 256)]:
    torch.manual_seed(0)
    m = pBLSTM(K, S).to(dev)
    m.precision = 'bf16'
    x = torch.randn(B, T, K, device=dev, requires_grad=True)
    lens = [T] * B
    for _ in range(2):
        out, _, _ = m(x, state_len=lens, pack_input=True)
        out.sum().backward()
    torch.cuda.synchronize()
    _lib.profile_read() if hasattr(_lib, 'profile_read') else None
    lib.ssasr_profile_enable(1)
    for _ in range(3):
        out, _, _ = m(x, state_len=lens, pack_input=True)
        out.sum().backward()
    prof = _lib.profile_read()
    lib.ssasr_profile_enable(0)
    msg = ', '.join('%s %.3f ms' % (k, v[0] / 3) for k, v in prof.items() if v[1] and k.startswith('rec'))
    print(f'{tag} B={B} T={T} K={K} S={S}: {msg}')
