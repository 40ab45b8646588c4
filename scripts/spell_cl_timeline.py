"""clock64 stamps of CTA 0 of the cluster decoder-step kernel (spell_cl.cu), averaged over the steps of one teacher-forced run:
0 h tile landed (MMA thread), 1 query accumulator ready (epilogue), 2 alpha image written (exchange thread), 3 all alpha rows
landed (MMA thread), 4 gate accumulator ready (epilogue), 5 h image written (exchange thread)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ss_asr_b200 import functional as Fk, _lib
from ss_asr_b200.asr import ASR
dev = 'cuda'
torch.manual_seed(1)
m = ASR(50, 256, 256, 128, 80, 0.9).to(dev).train()
import sys as _s
B, Tp, U = (int(_s.argv[1]) if len(_s.argv) > 1 else 256), 64, 41
g = torch.Generator().manual_seed(1)
enc = (0.3 * torch.randn(B, Tp, 512, generator=g)).to(dev)
lens = sorted([int(v) for v in torch.randint(48, 65, (B,), generator=g)], reverse=True)
tok = torch.randint(3, 50, (B, U), generator=g).to(torch.int32).to(dev)
lib = _lib.load()
Fk.set_cluster_speller(True)
for _ in range(2):
    m._spell(enc, lens, tok.clone(), [0] * U, 'bf16')
dbg = torch.zeros(U, 8, dtype=torch.int64, device=dev)
lib.ssasr_spell_cl_set_debug(dbg.data_ptr())
with torch.no_grad():
    pass
m._spell(enc.clone().requires_grad_(True), lens, tok.clone(), [0] * U, 'bf16')
torch.cuda.synchronize()
lib.ssasr_spell_cl_set_debug(None)
d = dbg.cpu()
names = ['h landed', 'q ready', 'alpha img', 'alpha landed', 'gates ready', 'h img']
print('step period (h landed -> h landed): %.0f cycles' % float((d[2:, 0] - d[1:-1, 0]).double().mean()))
for i in range(1, 6):
    print('%-14s +%.0f cycles after h landed' % (names[i], float((d[1:-1, i] - d[1:-1, 0]).double().mean())))
print('h img -> next h landed: %.0f' % float((d[2:, 0] - d[1:-1, 5]).double().mean()))

# backward kernel (layer-1 chain + attention): stamps of thread 128 of CTA 0
dbgb = torch.zeros(U, 8, dtype=torch.int64, device=dev)
e = enc.clone().requires_grad_(True)
logits, att, toks = m._spell(e, lens, tok.clone(), [0] * U, 'bf16')
lib.ssasr_spell_cl_set_debug_bwd(dbgb.data_ptr())
logits.backward(torch.ones_like(logits) * 1e-3)
torch.cuda.synchronize()
lib.ssasr_spell_cl_set_debug_bwd(None)
d = dbgb.cpu()
nb = ['dh landed', 'dG written', 'dalpha partials done', 'dalpha landed', 'dq written', 'dh accumulator ready', 'dh partials written']
sl = slice(2, U - 2)
print('backward step period: %.0f cycles' % float((d[3:U - 1, 0] - d[2:U - 2, 0]).double().mean()))
for i in range(1, 7):
    print('%-24s +%.0f cycles after dh landed' % (nb[i], float((d[sl, i] - d[sl, 0]).double().mean())))

# plain-recurrence mode (the layer-2 chain): forward stamps 0 (h landed), 4 (accumulator ready), 5 (h image written); backward
# stamps 0 (dh landed), 1 (dG written), 5 (dh accumulator ready), 6 (dh partials written)
lib.ssasr_spell_cl_set_debug_mode(1)
dbg.zero_(); dbgb.zero_()
lib.ssasr_spell_cl_set_debug(dbg.data_ptr()); lib.ssasr_spell_cl_set_debug_bwd(dbgb.data_ptr())
e = enc.clone().requires_grad_(True)
logits, att, toks = m._spell(e, lens, tok.clone(), [0] * U, 'bf16')
logits.backward(torch.ones_like(logits) * 1e-3)
torch.cuda.synchronize()
lib.ssasr_spell_cl_set_debug(None); lib.ssasr_spell_cl_set_debug_bwd(None); lib.ssasr_spell_cl_set_debug_mode(0)
d = dbg.cpu(); db = dbgb.cpu()
print('layer-2 forward step period %.0f cycles: accumulator ready +%.0f, h image +%.0f' % (
    float((d[3:U - 1, 0] - d[2:U - 2, 0]).double().mean()), float((d[sl, 4] - d[sl, 0]).double().mean()), float((d[sl, 5] - d[sl, 0]).double().mean())))
print('layer-2 backward step period %.0f cycles: dG written +%.0f, dh accumulator +%.0f, partials written +%.0f' % (
    float((db[3:U - 1, 0] - db[2:U - 2, 0]).double().mean()), float((db[sl, 1] - db[sl, 0]).double().mean()),
    float((db[sl, 5] - db[sl, 0]).double().mean()), float((db[sl, 6] - db[sl, 0]).double().mean())))
