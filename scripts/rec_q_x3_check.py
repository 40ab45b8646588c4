"""Exact-path (split-operand) quad-cluster recurrence `rec_q_fwd_kernel<R, false, true>` against the fp32 SIMT path and the
counter-barrier X3 kernels: max deviations and Listener time at decode shapes.  GPU only."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import las_oracle as O            # noqa: E402  (checker only)
from ss_asr_b200.asr import ASR              # noqa: E402

DEV = 'cuda'


def enc(m, x, lens, prec, env=None):
    if env is None:
        os.environ.pop('SSASR_REC_Q_X3', None)
    else:
        os.environ['SSASR_REC_Q_X3'] = env
    m.encoder.set_precision(prec)
    with torch.no_grad():
        e, _ = m.encoder(x, lens)
    torch.cuda.synchronize()
    m.encoder.set_precision('fp32')
    return e


for S, B, T in ((128, 37, 72), (256, 37, 72), (256, 70, 136), (256, 3, 40), (512, 21, 72), (512, 40, 136)):
    dims = (50, S, 32, 16, 24)
    sd = O.make_state_dict(*dims, seed=2)
    x, lens, _ = O.synth_batch(B, T, 24, 4, seed=5)
    m = ASR(*dims, 1.0).to(DEV).eval()
    m.load_state_dict(sd)
    for indep in (False, True):
        m.encoder.utterance_independent = indep
        ref = enc(m, x.to(DEV), lens, 'fp32')
        new32 = enc(m, x.to(DEV), lens, 'tf32x3')
        new16 = enc(m, x.to(DEV), lens, 'tf32x3', '16')
        old = enc(m, x.to(DEV), lens, 'tf32x3', '0')
        print('S=%d B=%d T=%d indep=%d: |new32-fp32| %.3g  |new16-fp32| %.3g  |old-fp32| %.3g  |new32-old| %.3g  nan %d' % (
            S, B, T, indep, float((new32 - ref).abs().max()), float((new16 - ref).abs().max()), float((old - ref).abs().max()),
            float((new32 - old).abs().max()), int(torch.isnan(new32).sum())), flush=True)

# decode-shaped timing: 512 utterances x 512 frames, default dims
dims = (50, 256, 256, 128, 80)
m = ASR(*dims, 1.0).to(DEV).eval()
m.encoder.utterance_independent = True
g = torch.Generator().manual_seed(0)
x = torch.randn(512, 512, 80, generator=g).to(DEV)
lens = sorted([int(v) for v in torch.randint(256, 513, (512,), generator=g)], reverse=True)
lens[0] = 512
for env in ('0', '16', None):
    enc(m, x, lens, 'tf32x3', env)
    t0 = time.perf_counter()
    for _ in range(3):
        enc(m, x, lens, 'tf32x3', env)
    print('Listener 512 x 512, tf32x3, SSASR_REC_Q_X3=%s: %.2f ms' % (env, (time.perf_counter() - t0) / 3 * 1e3), flush=True)
