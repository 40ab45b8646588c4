"""Per-phase clock64 timeline + CUDA-event timing of the recurrent BLSTM kernels, layer-1-like shape.
SSASR_REC_CLUSTER=0 selects the counter-barrier kernels (rec_tc.cu), default the cluster kernels (rec_cl.cu)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ss_asr_b200 import _lib
from ss_asr_b200.asr import pBLSTM
lib = _lib.load()
dev = 'cuda'
cluster = os.environ.get('SSASR_REC_CLUSTER', '1') != '0'
setdbg = lib.ssasr_rec_cl_set_debug if cluster else lib.ssasr_rec_tc_set_debug
origin = 1 if cluster else 0
for (B, T, K, S) in [(256, 256, 1024, 256)]:
    torch.manual_seed(0)
    m = pBLSTM(K, S).to(dev)
    m.precision = 'bf16'
    x = torch.randn(B, T, K, device=dev, requires_grad=True)
    lens = [T] * B
    dbg = torch.zeros(T, 12, dtype=torch.int64, device=dev)
    for which in ('fwd', 'bwd'):
        for _ in range(2):
            out, _, _ = m(x, state_len=lens, pack_input=True)
            out.sum().backward()
        torch.cuda.synchronize()
        lib.ssasr_profile_reset() if hasattr(lib, 'ssasr_profile_reset') else None
        out, _, _ = m(x, state_len=lens, pack_input=True)
        torch.cuda.synchronize()
        if which == 'fwd':
            setdbg(dbg.data_ptr())
            out, _, _ = m(x, state_len=lens, pack_input=True)
            torch.cuda.synchronize()
            setdbg(None)
        else:
            setdbg(dbg.data_ptr())
            lib.ssasr_rec_wide_set_debug(dbg.data_ptr())      # the K-split backward kernel (rec_wide.cu) stamps the same buffer
            out.sum().backward()
            torch.cuda.synchronize()
            setdbg(None)
            lib.ssasr_rec_wide_set_debug(None)
        d = dbg.cpu()
        print(f'B={B} T={T} K={K} S={S} cluster={cluster} {which}: stamps rel. to stamp[{origin}], steps 100..102')
        for s in range(100, 103):
            r = d[s]
            print('  step', s, [int(v - r[origin]) if v else None for v in r], ' next - this =', int(d[s + 1][origin] - r[origin]))
        lo, hi = T // 4, 3 * T // 4
        per = (d[hi][origin] - d[lo][origin]).item() / (hi - lo)
        print(f'  avg cycles/step {per:.0f}')
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out, _, _ = m(x, state_len=lens, pack_input=True)
        torch.cuda.synchronize()
        ev0.record()
        if which == 'fwd':
            for _ in range(5):
                out, _, _ = m(x, state_len=lens, pack_input=True)
        else:
            out.sum().backward()
        ev1.record()
        torch.cuda.synchronize()
        print(f'  {which} layer call (GEMM + recurrence): {ev0.elapsed_time(ev1) / (5 if which == "fwd" else 1):.3f} ms')
        dbg.zero_()
