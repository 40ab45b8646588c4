"""A/B of the recurrent backward kernels on one pBLSTM layer (layer-2 shape of C4): quad-cluster (rows 32 / 16) vs the 8-CTA
cluster kernel vs the counter-barrier kernel -- gradients must agree, CUDA-event time and clock64 timeline are printed."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ss_asr_b200 import _lib
from ss_asr_b200.asr import pBLSTM
lib = _lib.load()
dev = 'cuda'
shapes = [(256, 256, 1024, 256), (256, 128, 1024, 256), (70, 33, 40, 128)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for (B, T, K, S) in shapes:
    torch.manual_seed(0)
    m = pBLSTM(K, S).to(dev)
    m.precision = 'bf16'
    g = torch.Generator().manual_seed(5)
    lens = sorted([int(v) for v in torch.randint(3 * T // 4, T + 1, (B,), generator=g)], reverse=True)
    lens[0] = T
    x0 = torch.randn(B, T, K, generator=g)
    x0 = (x0 * (torch.arange(T)[None, :, None] < torch.tensor(lens)[:, None, None])).to(dev)
    gout = torch.randn(B, T // 2, 4 * S, generator=g).to(dev)
    res = {}
    for name, rows, cl in (('quad32', 32, 1), ('quad16', 16, 1), ('cl8', 0, 1), ('counter', 0, 0)):
        lib.ssasr_rec_q_set_rows(rows)
        lib.ssasr_rec_cl_enable(cl)
        ts = []
        for it in range(3):
            x = x0.clone().requires_grad_(True)
            m.zero_grad(set_to_none=True)
            out, _, _ = m(x, state_len=lens, pack_input=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if it == 2 and cl:
                dbg = torch.zeros(T, 12, dtype=torch.int64, device=dev)
                lib.ssasr_rec_cl_set_debug(dbg.data_ptr())
            e0.record()
            out.backward(gout)
            e1.record()
            torch.cuda.synchronize()
            lib.ssasr_rec_cl_set_debug(None)
            ts.append(e0.elapsed_time(e1))
        res[name] = (x.grad.clone(), {k: p.grad.clone() for k, p in m.named_parameters()})
        msg = '%-8s B=%d T=%d S=%d: layer backward %.3f ms (min of 3: %.3f)' % (name, B, T, S, ts[-1], min(ts))
        if cl:
            d = dbg.cpu()
            lo, hi = T // 4, 3 * T // 4
            per = (d[hi][1] - d[lo][1]).item() / (hi - lo)
            r = d[T // 2]
            msg += ' | %.0f cycles/step; stamps(step %d rel. to MMA start) %s' % (per, T // 2, [int(v - r[1]) if v else None for v in r[:12]])
        print(msg, flush=True)
    # forward: outputs of the four variants + timing / timeline of the recurrent forward kernel
    fres = {}
    for name, rows, cl in (('quad32', 32, 1), ('quad16', 16, 1), ('cl8', 0, 1), ('counter', 0, 0)):
        lib.ssasr_rec_q_set_rows(rows)
        lib.ssasr_rec_cl_enable(cl)
        with torch.no_grad():
            for it in range(3):
                if it == 2 and cl:
                    dbg = torch.zeros(T, 12, dtype=torch.int64, device=dev)
                    lib.ssasr_rec_cl_set_debug(dbg.data_ptr())
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out, _, _ = m(x0, state_len=lens, pack_input=True)
                e1.record()
                torch.cuda.synchronize()
                lib.ssasr_rec_cl_set_debug(None)
        fres[name] = out.clone()
        msg = '%-8s fwd B=%d T=%d S=%d: layer forward %.3f ms' % (name, B, T, S, e0.elapsed_time(e1))
        if cl:
            d = dbg.cpu()
            lo, hi = T // 4, 3 * T // 4
            per = (d[hi][1] - d[lo][1]).item() / (hi - lo)
            r = d[T // 2]
            msg += ' | %.0f cycles/step; stamps(step %d rel. to MMA start) %s' % (per, T // 2, [int(v - r[1]) if v else None for v in r[:12]])
        print(msg, flush=True)
    for name in ('quad32', 'quad16', 'cl8'):
        dmax = float((fres[name] - fres['counter']).abs().max())
        print('  fwd %s vs counter-barrier: max abs difference %.2e' % (name, dmax), flush=True)
        assert dmax < 2e-3, name
    ref = res['counter']
    for name in ('quad32', 'quad16', 'cl8'):
        worst = float((res[name][0] - ref[0]).norm() / ref[0].norm())
        for k in ref[1]:
            worst = max(worst, float((res[name][1][k] - ref[1][k]).norm() / (ref[1][k].norm() + 1e-20)))
        print('  %s vs counter-barrier: worst rel-L2 difference over dx and the 8 parameter gradients %.2e' % (name, worst), flush=True)
        assert worst < 2e-3, name
lib.ssasr_rec_q_set_rows(16)
lib.ssasr_rec_cl_enable(1)
print('ok')
