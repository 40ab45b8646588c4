"""Attend-and-spell loop alone (C4 shapes: B=256, T'=64, U=41): forward / backward time of the C call, layer-2 chain on the
second stream vs single stream."""
import os, sys, random, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ss_asr_b200 import functional as Fk
from ss_asr_b200.asr import ASR
dev = 'cuda'
torch.manual_seed(1)
m = ASR(50, 256, 256, 128, 80, 0.9).to(dev).train()
B, Tp, U = 256, 64, 41
g = torch.Generator().manual_seed(1)
enc = (0.1 * torch.randn(B, Tp, 512, generator=g)).to(dev).requires_grad_(True)
lens = sorted([int(v) for v in torch.randint(48, 65, (B,), generator=g)], reverse=True)
tok = torch.randint(3, 50, (B, U), generator=g).to(torch.int32).to(dev)
for dual in (False, True, False, True):
    Fk.set_dual_stream_speller(dual)
    res = []
    for it in range(6):
        random.seed(it)
        modes = [0 if random.random() <= 0.9 else 2 for _ in range(U)]
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        logits, att, toks = m._spell(enc, lens, tok, modes, 'bf16')
        e[1].record()
        logits.backward(torch.ones_like(logits) * 1e-3)
        e[2].record()
        torch.cuda.synchronize()
        res.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    res = res[2:]
    print('dual' if dual else 'single', 'fwd %.3f ms  bwd %.3f ms' % (sum(r[0] for r in res) / len(res), sum(r[1] for r in res) / len(res)))
