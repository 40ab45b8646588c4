"""2+ GPU check of the data-parallel path with deferred weight gradients: GradSync's bucketed all-reduce (launched from the
post-accumulate hooks, on the side stream) must give the same averaged gradients as a plain all-reduce after an in-order
backward.  torchrun --nproc-per-node N scripts/dp_check.py"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import las_oracle as O
from ss_asr_b200 import functional as Fk
from ss_asr_b200.asr import ASR
from ss_asr_b200.functional import asr_loss
from ss_asr_b200.parallel import GradSync

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
dims = (50, 64, 64, 32, 40)
sd = O.make_state_dict(*dims, seed=1)
x, lens, y = O.synth_batch(70, 64, dims[4], 6, seed=100 + rank)
res = {}
for on in (False, True):
    Fk.set_overlap_wgrad(on)
    m = ASR(*dims, 1.0).to(dev)
    m.load_state_dict(sd)
    m.train_precision = 'bf16'
    m.train()
    if on:
        sync = GradSync(m, world)
        for _ in range(3):
            m.zero_grad(set_to_none=True)
            _, logits, _ = m(x.to(dev), 7, teacher=y.to(dev), state_len=lens)
            sync.backward(asr_loss(logits, y.to(dev)))
        sync.close()
    else:
        _, logits, _ = m(x.to(dev), 7, teacher=y.to(dev), state_len=lens)
        asr_loss(logits, y.to(dev)).backward()
        for p in m.parameters():
            dist.all_reduce(p.grad)
            p.grad /= world
    torch.cuda.synchronize()
    res[on] = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
Fk.set_overlap_wgrad(False)
worst = 0.0
for k in res[False]:
    a, b = res[True][k], res[False][k]
    rel = float((a - b).norm()) / (float(b.norm()) + 1e-12)
    worst = max(worst, rel)
    assert rel < 1e-4 or float((a - b).norm()) < 1e-7, (k, rel)
print('rank %d: DP gradients with deferred weight gradients match (worst rel diff %.2e)' % (rank, worst))
dist.destroy_process_group()
