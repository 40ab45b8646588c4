"""Quick on-GPU numerical check of every kernel family against the CPU oracle; prints max diffs."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fbank_oracle as FB  # noqa: E402
from oracle import las_oracle as O  # noqa: E402
from ss_asr_b200 import _lib, functional as Fk  # noqa: E402
from ss_asr_b200.asr import ASR  # noqa: E402
from ss_asr_b200 import preprocess as PP  # noqa: E402

dev = 'cuda'


def gemm_check():
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    for (M, N, K, akm, bkm) in [(70, 50, 33, 1, 1), (128, 64, 80, 1, 0), (65, 130, 257, 0, 0), (256, 1024, 1024, 1, 1),
                                (40, 24, 500, 0, 1)]:
        A = torch.randn(M, K, generator=g)
        B = torch.randn(K, N, generator=g)
        bias = torch.randn(N, generator=g)
        ref = A.double() @ B.double() + bias.double()
        Ad = (A if akm else A.t().contiguous()).to(dev)
        Bd = (B.t().contiguous() if bkm else B).to(dev)
        Cd = torch.empty(M, N, device=dev)
        _lib.check(lib.ssasr_gemm_f32(M, N, K, Ad.data_ptr(), K if akm else M, akm, Bd.data_ptr(), K if bkm else N, bkm,
                                      Cd.data_ptr(), N, bias.to(dev).data_ptr(), 0, 0, _lib.stream()), 'gemm')
        print('gemm', (M, N, K, akm, bkm), 'max err', float((Cd.cpu().double() - ref).abs().max()))


def tc_check():
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    for (M, N, K, ak, bk) in [(128, 128, 64, 0, 0), (256, 256, 256, 0, 0), (300, 200, 80, 0, 0), (1024, 2048, 1024, 0, 0),
                              (128, 128, 512, 16, 0), (256, 128, 1000, 0, 8), (2048, 80, 4096, 0, 0)]:
        Kt = K + max(ak, bk)
        Kp = (Kt + 7) // 8 * 8
        A = torch.randn(M, Kp, generator=g).to(dev)
        B = torch.randn(N, Kp, generator=g).to(dev)
        bias = torch.randn(N, generator=g).to(dev)
        Ab, Bb = A.to(torch.bfloat16), B.to(torch.bfloat16)
        ref = Ab[:, ak:ak + K].double() @ Bb[:, bk:bk + K].double().t() + bias.double()
        C = torch.zeros(M, N, device=dev)
        _lib.check(lib.ssasr_gemm_bf16_tc(M, N, K, Ab.data_ptr(), Kp, ak, Bb.data_ptr(), Kp, bk, C.data_ptr(), N,
                                          bias.data_ptr(), 0, _lib.stream()), 'gemm_tc')
        torch.cuda.synchronize()
        print('gemm_tc', (M, N, K, ak, bk), 'max err', float((C.double() - ref).abs().max()), 'ref max', float(ref.abs().max()))
    # throughput
    M, N, K = 131072, 2048, 1024
    Ab = torch.randn(M, K, device=dev).to(torch.bfloat16)
    Bb = torch.randn(N, K, device=dev).to(torch.bfloat16)
    C = torch.empty(M, N, device=dev)
    for _ in range(2):
        lib.ssasr_gemm_bf16_tc(M, N, K, Ab.data_ptr(), K, 0, Bb.data_ptr(), K, 0, C.data_ptr(), N, None, 0, _lib.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lib.ssasr_gemm_bf16_tc(M, N, K, Ab.data_ptr(), K, 0, Bb.data_ptr(), K, 0, C.data_ptr(), N, None, 0, _lib.stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print('gemm_tc %dx%dx%d: %.3f ms, %.1f TFLOP/s' % (M, N, K, ms, 2.0 * M * N * K / ms / 1e9))


def tf32_check():
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    for (M, N, K) in [(128, 128, 32), (256, 256, 256), (300, 200, 80), (1024, 2048, 1024), (4096, 2048, 1024)]:
        A = torch.randn(M, K, generator=g).to(dev)
        B = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
        bias = torch.randn(N, generator=g).to(dev)
        ref = A.double() @ B.double().t() + bias.double()
        C = torch.zeros(M, N, device=dev)
        Al, Bl = torch.empty(2 * M * K, device=dev), torch.empty(2 * N * K, device=dev)
        _lib.check(lib.ssasr_gemm_tf32x3(M, N, K, A.data_ptr(), Al.data_ptr(), K, B.data_ptr(), Bl.data_ptr(), K, C.data_ptr(), N,
                                         bias.data_ptr(), 0, _lib.stream()), 'tf32x3')
        torch.cuda.synchronize()
        f32 = (A @ B.t() + bias)
        print('tf32x3', (M, N, K), 'max err vs fp64', float((C.double() - ref).abs().max()), ' torch fp32 err', float((f32.double() - ref).abs().max()))
    M, N, K = 131072, 2048, 1024
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); C = torch.empty(M, N, device=dev)
    Al, Bl = torch.empty(2 * M * K, device=dev), torch.empty(2 * N * K, device=dev)
    for _ in range(2):
        lib.ssasr_gemm_tf32x3(M, N, K, A.data_ptr(), Al.data_ptr(), K, B.data_ptr(), Bl.data_ptr(), K, C.data_ptr(), N, None, 0, _lib.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        lib.ssasr_gemm_tf32x3(M, N, K, A.data_ptr(), Al.data_ptr(), K, B.data_ptr(), Bl.data_ptr(), K, C.data_ptr(), N, None, 0, _lib.stream())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('tf32x3 %dx%dx%d: %.3f ms, %.1f effective TFLOP/s' % (M, N, K, ms, 2.0 * M * N * K / ms / 1e9))


def tn_check():
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    for (M, N, K, ak, bk) in [(128, 128, 64, 0, 0), (128, 128, 128, 0, 0), (256, 256, 512, 0, 0), (2048, 80, 4096, 0, 0),
                              (1024, 256, 1000, 1, 0), (1024, 256, 1000, 0, 3), (50, 256, 333, 0, 0)]:
        Kt = K + max(ak, bk)
        Mp, Np = (M + 7) // 8 * 8, (N + 7) // 8 * 8
        A = torch.randn(Kt, Mp, generator=g).to(dev).to(torch.bfloat16)
        B = torch.randn(Kt, Np, generator=g).to(dev).to(torch.bfloat16)
        ref = A[ak:ak + K, :M].double().t() @ B[bk:bk + K, :N].double()
        C = torch.zeros(M, N, device=dev)
        _lib.check(lib.ssasr_gemm_bf16_tc_tn(M, N, K, A.data_ptr(), Mp, ak, B.data_ptr(), Np, bk, C.data_ptr(), N, 0, _lib.stream()),
                   'gemm_tc_tn')
        torch.cuda.synchronize()
        print('gemm_tc_tn', (M, N, K, ak, bk), 'max err', float((C.double() - ref).abs().max()), 'ref max', float(ref.abs().max()))


def bf16_check(dims=(50, 256, 256, 128, 80), B=8, T=128, U=20):
    sd = O.make_state_dict(*dims, seed=1)
    x, lens, y = O.synth_batch(B, T, dims[4], U, seed=1234)
    loss_o, logits_o, att_o, enc_o, grads_o = O.train_step_grads(sd, x, lens, y)
    m = ASR(*dims, 1.0).to(dev)
    m.load_state_dict(sd)
    m.train_precision = 'bf16'
    m.train()
    ans_len = logits_o.shape[1]
    el, logits, att = m(x.to(dev), ans_len, teacher=y.to(dev), state_len=lens)
    print('bf16 logits max err', float((logits.detach().cpu() - logits_o).abs().max()), 'att', float((att - att_o).abs().max()))
    loss = Fk.asr_loss(logits, y.to(dev))
    print('bf16 loss', float(loss), float(loss_o))
    loss.backward()
    worst = 0
    for k, p in m.named_parameters():
        a, b = p.grad.cpu().double(), grads_o[k].double()
        rel = float((a - b).norm() / (b.norm() + 1e-12))
        cos = float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
        worst = max(worst, rel)
        if rel > 1e-2:
            print('  grad', k, 'rel', rel, 'cos', cos)
    print('bf16 worst grad rel-L2', worst)


def fbank_check():
    z = np.load(os.path.join(ROOT, 'tests/golden/fbank_1s.npz'))
    for n_mels, key in ((80, 'fb16_80'), (40, 'fb16_40')):
        out = PP.log_fbank_batch([z['y16']], 16000, n_mels)[0]
        print('fbank 16k', n_mels, out.shape, 'max err', float(np.abs(out - z[key]).max()))
    out = PP.log_fbank_batch([z['y16'][:11025]], 22050, 40)[0]
    print('fbank 22k', out.shape, 'max err', float(np.abs(out - z['fb22_40']).max()))
    rng = np.random.RandomState(3)
    ys = [(0.1 * rng.randn(n)).astype(np.float32) for n in (16000, 8000, 12345, 401, 3200)]
    outs = PP.log_fbank_batch(ys, 16000, 80)
    for y, o in zip(ys, outs):
        w = FB.log_fbank(y, 16000, 80)
        print('fbank ragged', len(y), o.shape, w.shape, 'max err', float(np.abs(o - w).max()))


def las_check(dims, B, T, U, lens=None, seed=1234):
    sd = O.make_state_dict(*dims, seed=1)
    if lens is None:
        x, lens, y = O.synth_batch(B, T, dims[4], U, seed=seed)
    else:
        g = torch.Generator().manual_seed(99)
        x = torch.randn(B, T, dims[4], generator=g)
        for i, l in enumerate(lens):
            x[i, l:] = 0
        _, _, y = O.synth_batch(B, T, dims[4], U, seed=5)
    t0 = time.time()
    loss_o, logits_o, att_o, enc_o, grads_o = O.train_step_grads(sd, x, lens, y)
    print('oracle %.1fs' % (time.time() - t0))
    m = ASR(*dims, 1.0).to(dev)
    m.load_state_dict(sd)
    ans_len = int(max((y != 0).sum(-1) + 1)) - 1
    enc, el = m.encoder(x.to(dev), lens)
    print('enc shape', tuple(enc.shape), 'lens', el, 'max err', float((enc.cpu() - enc_o).abs().max()))
    el, logits, att = m(x.to(dev), ans_len, teacher=y.to(dev), state_len=lens)
    print('logits max err', float((logits.cpu() - logits_o).abs().max()), 'att max err', float((att - att_o).abs().max()))
    loss = Fk.asr_loss(logits, y.to(dev))
    print('loss', float(loss), float(loss_o))
    loss.backward()
    gtot = float(torch.sqrt(sum(v.double().pow(2).sum() for v in grads_o.values())))
    worst = 0
    for k, p in m.named_parameters():
        d = float((p.grad.cpu().double() - grads_o[k].double()).norm())
        rel = d / (float(grads_o[k].double().norm()) + 1e-6 * gtot)
        worst = max(worst, rel)
        if rel > 1e-4:
            print('  grad', k, 'rel', rel)
    print('worst grad rel-L2', worst)
    # greedy
    with torch.no_grad():
        el2, gl, _ = m(x.to(dev), U + 3, state_len=lens)
        _, gl_o, _, _ = O.asr_forward(sd, x, lens, U + 3)
    print('greedy argmax equal', bool((gl.argmax(-1).cpu() == gl_o.argmax(-1)).all()), 'max err',
          float((gl.cpu() - gl_o).abs().max()))


def decode_check():
    z = np.load(os.path.join(ROOT, 'tests/golden/decode_default.npz'))
    sd = O.make_state_dict(50, 256, 256, 128, 80, seed=1)
    sd['char_trans.weight'] = sd['char_trans.weight'] * 20.0
    m = ASR(50, 256, 256, 128, 80, 1.0).to(dev)
    m.load_state_dict(sd)
    Ts = [int(v) for v in z['Ts']]
    xs = [torch.randn(1, Ti, 80, generator=torch.Generator().manual_seed(7000 + i)) for i, Ti in enumerate(Ts)]
    order = sorted(range(len(Ts)), key=lambda i: -Ts[i])
    Tm = max(Ts)
    xb = torch.zeros(len(Ts), Tm, 80)
    for j, i in enumerate(order):
        xb[j, :Ts[i]] = xs[i][0]
    ids = m.decode_batch(xb.to(dev), [Ts[i] for i in order])
    for j, i in enumerate(order):
        got = O.ids_to_str(ids[j])
        print('decode', i, Ts[i], got == str(z['margin_lm00'][i]), got[:30])


if __name__ == '__main__':
    torch.manual_seed(0)
    which = sys.argv[1:] or ['gemm', 'fbank', 'tiny', 'default', 'decode']
    if 'gemm' in which:
        gemm_check()
    if 'tc' in which:
        tc_check()
    if 'tf32' in which:
        tf32_check()
    if 'tn' in which:
        tn_check()
    if 'bf16' in which:
        bf16_check()
        bf16_check((50, 32, 48, 16, 20), 7, 64, 9)
    if 'fbank' in which:
        fbank_check()
    if 'tiny' in which:
        las_check((50, 16, 16, 8, 12), 5, 77, 10, lens=[61, 53, 40, 33, 9])
        las_check((50, 32, 48, 16, 20), 7, 64, 9)
    if 'default' in which:
        las_check((50, 256, 256, 128, 80), 8, 128, 20)
    if 'decode' in which:
        decode_check()
    torch.cuda.synchronize()
    print('done')
