"""C5 (long utterances, 512-dim BLSTM): parity of a small batch against the oracle + timing of the B=32 step."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import las_oracle as O
from ss_asr_b200.asr import ASR
from ss_asr_b200.functional import asr_loss
dev = 'cuda'
dims = (50, 512, 256, 128, 80)
sd = O.make_state_dict(*dims, seed=1)
x, lens, y = O.synth_batch(2, 1600, 80, 20, seed=1234)
t0 = time.time()
loss_o, logits_o, att_o, enc_o, grads_o = O.train_step_grads(sd, x, lens, y)
print('oracle B=2 T=1600 S=512: %.1fs' % (time.time() - t0))
for prec in ('fp32', 'bf16'):
    m = ASR(*dims, 1.0).to(dev)
    m.load_state_dict(sd)
    m.train_precision = prec
    m.train()
    _, logits, att = m(x.to(dev), logits_o.shape[1], teacher=y.to(dev), state_len=lens)
    loss = asr_loss(logits, y.to(dev))
    loss.backward()
    worst = max(float((p.grad.cpu().double() - grads_o[k].double()).norm() / (grads_o[k].double().norm() + 1e-12))
                for k, p in m.named_parameters() if float(grads_o[k].norm()) > 1e-6)
    print(prec, 'logits err', float((logits.detach().cpu() - logits_o).abs().max()), 'loss', float(loss), float(loss_o),
          'worst grad rel', worst)
# timing, B=32
x, lens, y = O.synth_batch(32, 1600, 80, 100, seed=1234)
m = ASR(*dims, 0.9).to(dev)
m.train_precision = 'bf16'
m.train()
opt = torch.optim.Adadelta(m.parameters(), lr=1.0, eps=1e-8)
xd, yd = x.to(dev), y.to(dev)
ans = int(max((y != 0).sum(-1) + 1)) - 1
def step():
    opt.zero_grad(set_to_none=True)
    _, logits, _ = m(xd, ans, teacher=yd, state_len=lens)
    loss = asr_loss(logits, yd)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
    opt.step()
for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print('C5 train step B=32 T=1600 S_enc=512 U=100: %.1f ms = %.1f utt/s (reference CPU, survey: 0.129 utt/s)' % (ms, 32 / ms * 1e3))
m.eval()
t0 = time.time()
ids = m.decode_batch(xd[:8], lens[:8])
torch.cuda.synchronize()
print('C5 decode 8 utts: %.3fs' % (time.time() - t0), [len(i) for i in ids])
