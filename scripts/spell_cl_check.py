"""Cluster-persistent decoder-step kernel (spell_cl.cu) against the per-step kernels on the same inputs (C4 decoder shapes),
then the forward / backward time of the attend-and-spell call with each.  usage: python scripts/spell_cl_check.py [B]"""
import os, sys, random, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ss_asr_b200 import functional as Fk
from ss_asr_b200.asr import ASR
dev = 'cuda'
torch.manual_seed(1)
m = ASR(50, 256, 256, 128, 80, 0.9).to(dev).train()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
CLB = (sys.argv[2] != '0') if len(sys.argv) > 2 else True
Tp, U = 64, 41
g = torch.Generator().manual_seed(1)
enc0 = (0.3 * torch.randn(B, Tp, 512, generator=g)).to(dev)
lens = sorted([int(v) for v in torch.randint(48, 65, (B,), generator=g)], reverse=True)
for i, l in enumerate(lens):
    enc0[i, l:] = 0
tok = torch.randint(3, 50, (B, U), generator=g).to(torch.int32).to(dev)
gl = torch.randn(B, U, 50, generator=g).to(dev) * 1e-2


def run(cl, modes, clb=None):
    Fk.set_cluster_speller(cl, clb)
    enc = enc0.clone().requires_grad_(True)
    m.zero_grad(set_to_none=True)
    logits, att, toks = m._spell(enc, lens, tok.clone(), modes, 'bf16')
    (logits * gl).sum().backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    return logits.detach(), att.detach(), enc.grad.clone(), grads


ref = run(False, [0] * U)
new = run(True, [0] * U, CLB)
print('logits max|d| %.3e (max|ref| %.3e)' % (float((ref[0] - new[0]).abs().max()), float(ref[0].abs().max())))
print('att    max|d| %.3e (max ref %.3e)' % (float((ref[1] - new[1]).abs().max()), float(ref[1].max())))
print('denc   rel-L2 %.3e' % (float((ref[2] - new[2]).norm() / ref[2].norm())))
for k in ref[3]:
    a, b = new[3][k], ref[3][k]
    print('  grad %-36s rel-L2 %.3e' % (k, float((a - b).norm() / (b.norm() + 1e-30))))
print('nan check', bool(torch.isnan(new[0]).any()), bool(torch.isnan(new[1]).any()))

for cl in (False, True, False, True):
    Fk.set_cluster_speller(cl, cl and CLB)
    res = []
    for it in range(6):
        random.seed(it)
        modes = [0 if random.random() <= 0.9 else 2 for _ in range(U)]
        enc = enc0.clone().requires_grad_(True)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        logits, att, toks = m._spell(enc, lens, tok.clone(), modes, 'bf16')
        e[1].record()
        logits.backward(gl)
        e[2].record()
        torch.cuda.synchronize()
        res.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    res = res[2:]
    print('cluster' if cl else 'per-step', 'fwd %.3f ms  bwd %.3f ms' % (sum(r[0] for r in res) / len(res), sum(r[1] for r in res) / len(res)))

# per-family device time of one forward call on each path (kernels timed one launch at a time)
from ss_asr_b200 import _lib
lib = _lib.load()
for cl in (False, True):
    Fk.set_cluster_speller(cl, cl and CLB)
    random.seed(3)
    modes = [0 if random.random() <= 0.9 else 2 for _ in range(U)]
    lib.ssasr_profile_enable(1)
    _lib.profile_read()
    import time
    t0 = time.perf_counter()
    with torch.no_grad():
        pass
    enc = enc0.clone().requires_grad_(True)
    logits, att, toks = m._spell(enc, lens, tok.clone(), modes, 'bf16')
    t1 = time.perf_counter()
    fam = _lib.profile_read()
    logits.backward(gl)
    famb = _lib.profile_read()
    lib.ssasr_profile_enable(0)
    print('cluster' if cl else 'per-step', 'host %.2f ms' % ((t1 - t0) * 1e3), 'fwd:', {k: (round(v[0], 3), v[1]) for k, v in fam.items() if v[1]},
          'bwd:', {k: (round(v[0], 3), v[1]) for k, v in famb.items() if v[1]})
