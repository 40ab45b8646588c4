"""Host-side enqueue time of one C4 train step (is the step host-bound?): wall time to ENQUEUE n steps (no sync) vs the
device time of the same steps; optional cProfile of the enqueue loop.  usage: python scripts/host_time.py [--prof]"""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as BN
from ss_asr_b200 import functional as Fk
from ss_asr_b200.functional import asr_loss
from ss_asr_b200.optim import FusedAdadelta

dev = torch.device('cuda', 0)
m = BN.fresh_model(dev); m.train_precision = 'bf16'; m.train(); m.att_on_device = True
opt = FusedAdadelta(m.parameters(), lr=1.0, eps=1e-8)
Fk.set_overlap_wgrad(True)
x, lens, y = BN.synth_batch(256, 512, 80, 40)
xd, yd = x.to(dev), y.to(dev)
ans = int(max((y != 0).sum(-1) + 1)) - 1
tm = {'fwd': 0.0, 'bwd': 0.0, 'opt': 0.0}
def step(rec=False):
    t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    _, logits, _ = m(xd, ans, teacher=yd, state_len=lens)
    loss = asr_loss(logits, yd)
    t1 = time.perf_counter()
    loss.backward()
    t2 = time.perf_counter()
    opt.step_clipped(5.0)
    t3 = time.perf_counter()
    if rec:
        tm['fwd'] += t1 - t0; tm['bwd'] += t2 - t1; tm['opt'] += t3 - t2
for _ in range(5): step()
torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(n): step(True)
e1.record(); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print('enqueue %.2f ms/step (fwd %.2f bwd %.2f opt %.2f), wall incl. drain %.2f ms/step, device %.2f ms/step'
      % ((t1 - t0) / n * 1e3, tm['fwd'] / n * 1e3, tm['bwd'] / n * 1e3, tm['opt'] / n * 1e3, (t2 - t0) / n * 1e3, e0.elapsed_time(e1) / n))
# one step with a sync in the middle so that the queue is empty: pure host cost without back-pressure
torch.cuda.synchronize()
for k in tm: tm[k] = 0.0
for _ in range(5):
    torch.cuda.synchronize(); step(True)
print('host cost with an empty queue: fwd %.2f bwd %.2f opt %.2f ms' % (tm['fwd'] / 5 * 1e3, tm['bwd'] / 5 * 1e3, tm['opt'] / 5 * 1e3))
if '--prof' in sys.argv:
    pr = cProfile.Profile(); torch.cuda.synchronize(); pr.enable()
    for _ in range(5): step()
    pr.disable(); torch.cuda.synchronize()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45); print(s.getvalue()[:9000])
