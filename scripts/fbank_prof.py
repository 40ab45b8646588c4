import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ss_asr_b200 import preprocess as PP
n_utt, n = 1024, 160000
audio = 0.1 * torch.randn(n_utt * n, device="cuda")
off = [i * n for i in range(n_utt + 1)]
plan = PP.FbankPlan(off, 16000, 80)
fb = plan.run(audio)
for _ in range(3):
    plan.run(audio, out=fb)
torch.cuda.synchronize()
print('ok')
