"""Parity of the CUDA path (through the C ABI, libssasr.so) against the CPU oracle and the committed golden
vectors of the unmodified reference.  Needs a B200: `pytest -m gpu`.

Tolerances (fp32 path): logits/attention atol 1e-5, loss rel 1e-6 (+1e-6 abs), gradients
rel-L2 1e-4 with an absolute floor of 1e-6 * global grad norm, fbank atol 1e-4 in the log domain, greedy
token sequences identical."""
import os

import numpy as np
import pytest
import torch

from oracle import fbank_oracle as FB
from oracle import las_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _model(dims, sd, tf=1.0):
    from ss_asr_b200.asr import ASR
    m = ASR(*dims, tf).to(DEV)
    m.load_state_dict(sd)
    return m


class _CharLM(torch.nn.Module):
    """Same parameter structure as the reference CharLM (charlm.py:5-44); only its state_dict is consumed."""

    def __init__(self, n=50, h=128):
        super().__init__()
        self.emb = torch.nn.Embedding(n, h)
        self.layer_1 = torch.nn.GRUCell(h, h)
        self.layer_2 = torch.nn.GRUCell(h, h)
        self.out = torch.nn.Linear(h, n)
        self.hidden_size = h

    def forward(self, x, h_1, h_2):                    # charlm.py:46-57
        x = self.emb(x.long())
        h_1 = self.layer_1(x, h_1)
        h_2 = self.layer_2(h_1, h_2)
        return self.out(h_2), (h_1, h_2)

    def init_hidden(self, batch_size, device):         # charlm.py:59-61
        return (torch.zeros(batch_size, self.hidden_size).to(device), torch.zeros(batch_size, self.hidden_size).to(device))


def _charlm_module(lm_sd):
    lm = _CharLM()
    lm.load_state_dict(lm_sd)
    return lm.to(DEV).eval()


def _check_grads(model, ref_grads, rel=1e-4):
    gtot = float(torch.sqrt(sum(torch.as_tensor(v).double().pow(2).sum() for v in ref_grads.values())))
    for k, p in model.named_parameters():
        want = torch.as_tensor(ref_grads[k]).double()
        d = float((p.grad.cpu().double() - want).norm())
        assert d <= rel * float(want.norm()) + 1e-6 * gtot, (k, d, float(want.norm()))


# ------------------------------------------------------------------------------------------------ library
def test_library_loaded_is_in_tree():
    from ss_asr_b200 import _lib
    _lib.load()
    assert os.path.samefile(_lib.lib_path(), os.path.join(os.path.dirname(os.path.dirname(__file__)), 'ss_asr_b200',
                                                          'libssasr.so'))
    with open('/proc/self/maps') as f:
        assert 'libssasr.so' in f.read()


def test_cpu_tensors_fail_loudly():
    sd = O.make_state_dict(50, 16, 16, 8, 12, seed=1)
    from ss_asr_b200.asr import ASR
    m = ASR(50, 16, 16, 8, 12, 1.0)
    m.load_state_dict(sd)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m(torch.randn(2, 16, 12), 3, state_len=[16, 16])


@pytest.mark.parametrize('M,N,K,akm,bkm', [(70, 50, 33, 1, 1), (128, 64, 80, 1, 0), (65, 130, 257, 0, 0), (1, 7, 5, 1, 1),
                                           (40, 24, 500, 0, 1)])
def test_gemm_f32(M, N, K, akm, bkm):
    from ss_asr_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(M * 7 + N)
    A, B, bias = torch.randn(M, K, generator=g), torch.randn(K, N, generator=g), torch.randn(N, generator=g)
    ref = (A.double() @ B.double() + bias.double())
    Ad = (A if akm else A.t().contiguous()).to(DEV)
    Bd = (B.t().contiguous() if bkm else B).to(DEV)
    Cd = torch.empty(M, N, device=DEV)
    _lib.check(lib.ssasr_gemm_f32(M, N, K, Ad.data_ptr(), K if akm else M, akm, Bd.data_ptr(), K if bkm else N, bkm,
                                  Cd.data_ptr(), N, bias.to(DEV).data_ptr(), 0, 0, _lib.stream()), 'gemm')
    assert float((Cd.cpu().double() - ref).abs().max()) < 1e-5 * K ** 0.5 * 4


# ------------------------------------------------------------------------------------------------ fbank
def test_fbank_golden(golden_dir):
    from ss_asr_b200 import preprocess as PP
    z = np.load(os.path.join(golden_dir, 'fbank_1s.npz'))
    assert np.abs(PP.log_fbank_batch([z['y16']], 16000, 80)[0] - z['fb16_80']).max() < 1e-4
    assert np.abs(PP.log_fbank_batch([z['y16']], 16000, 40)[0] - z['fb16_40']).max() < 1e-4
    assert np.abs(PP.log_fbank(z['y16'][:11025], 22050) - z['fb22_40']).max() < 1e-4     # reference defaults


def test_fbank_ragged_batch_and_edges():
    from ss_asr_b200 import preprocess as PP
    rng = np.random.RandomState(3)
    ns = [16000, 8000, 12345, 401, 201, 3200, 160000, 1599, 1600, 1601]
    ys = [(0.1 * rng.randn(n)).astype(np.float32) for n in ns]
    ys[3][:] = 0.0                                                   # digital silence -> log(eps)
    outs = PP.log_fbank_batch(ys, 16000, 80)
    for y, o in zip(ys, outs):
        w = FB.log_fbank(y, 16000, 80)
        assert o.shape == w.shape and o.dtype == np.float32
        assert np.abs(o - w).max() < 1e-4
    assert np.allclose(outs[3], np.log(np.finfo(float).eps))
    with pytest.raises(ValueError):
        PP.log_fbank_batch([np.zeros(200, np.float32)], 16000, 80)  # numpy reflect pad needs > ws//2 samples


def test_fbank_host_pipeline_matches_device_path():
    """FbankPlan.run_host (pinned host buffers in and out, three streams, groups of utterances) == FbankPlan.run on the same
    ragged batch, twice in a row (buffer reuse across calls)."""
    from ss_asr_b200 import preprocess as PP
    rs = np.random.RandomState(5)
    lens = [int(v) for v in rs.randint(3000, 40000, size=37)]
    off = [0]
    for n in lens:
        off.append(off[-1] + n)
    plan = PP.FbankPlan(off, 16000, 80, device=DEV)
    for rep in range(2):
        host = torch.from_numpy((0.1 * rs.randn(off[-1])).astype(np.float32)).pin_memory()
        want = plan.run(host.to(DEV)).cpu()
        out = torch.empty(want.shape, dtype=torch.float32).pin_memory()
        plan.run_host(host, out, n_parts=5)
        plan.sync()
        assert torch.equal(out, want)


def test_fbank_to_listener_handoff():
    """fbank output -> padded, length-sorted Listener batch on the device == the reference's disk format read back
    (zero_pad of preprocess.py:253-269 + the prepare_x length count)."""
    from ss_asr_b200 import preprocess as PP
    from ss_asr_b200.dataset import prepare_x
    rs = np.random.RandomState(2)
    ys = [(0.1 * rs.randn(n)).astype(np.float32) for n in (9000, 20000, 4100, 15000)]
    off = [0]
    for y in ys:
        off.append(off[-1] + len(y))
    fb, foff = PP.log_fbank_device(torch.from_numpy(np.concatenate(ys)).to(DEV), off, 16000, 40)
    x, lens, order = PP.fbank_to_listener_batch(fb, foff)
    assert order == [1, 3, 0, 2] and lens == sorted(lens, reverse=True) and x.shape == (4, lens[0], 40)
    padded = np.zeros((4, lens[0], 40))                             # what preprocess.py would have written, per utterance
    for row, i in enumerate(order):
        padded[row, :lens[row]] = fb[foff[i]:foff[i + 1]].cpu().numpy()
    x_ref, lens_ref = prepare_x(torch.from_numpy(padded).unsqueeze(0), DEV)
    assert lens_ref == lens and torch.equal(x_ref, x)


def test_fbank_full_size_properties():
    """C2 shape (160 000-sample utterances): linearity in the power domain and agreement with the oracle on
    sampled utterances."""
    from ss_asr_b200 import preprocess as PP
    g = torch.Generator().manual_seed(1234)
    n_utt, n = 64, 160000
    audio = (0.1 * torch.randn(n_utt * n, generator=g)).to(DEV)
    off = [i * n for i in range(n_utt + 1)]
    out, foff = PP.log_fbank_device(audio, off, 16000, 80)
    assert out.shape == (n_utt * 1001, 80) and foff[1] == 1001
    out2, _ = PP.log_fbank_device(audio * 2.0, off, 16000, 80)
    assert float((out2 - out - np.log(4.0)).abs().max()) < 1e-4      # power scales by 4 -> log shifts by ln 4
    for u in (0, 17, 63):
        w = FB.log_fbank(audio[u * n:(u + 1) * n].cpu().numpy(), 16000, 80)
        assert np.abs(out[u * 1001:(u + 1) * 1001].cpu().numpy() - w).max() < 1e-4


# ------------------------------------------------------------------------------------------------ LAS
def test_tiny_golden_forward_loss_grads(golden_dir):
    from ss_asr_b200.functional import asr_loss
    z = np.load(os.path.join(golden_dir, 'las_tiny.npz'))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd.')}
    dims = tuple(int(v) for v in z['dims'])
    x, lens, y = torch.from_numpy(z['x']), [int(v) for v in z['lens']], torch.from_numpy(z['y'])
    m = _model(dims, sd)
    enc, el = m.encoder(x.to(DEV), lens)
    assert el == [int(v) for v in z['enc_len']] and tuple(enc.shape) == z['enc'].shape
    assert np.abs(enc.detach().cpu().numpy() - z['enc']).max() < 1e-5
    U = z['logits'].shape[1]
    el, logits, att = m(x.to(DEV), U, teacher=y.to(DEV), state_len=lens)
    assert not att.is_cuda and logits.is_cuda                        # reference return convention (asr.py:103,110)
    assert np.abs(logits.detach().cpu().numpy() - z['logits']).max() < 1e-5
    assert np.abs(att.numpy() - z['att']).max() < 1e-5
    loss = asr_loss(logits, y.to(DEV))
    assert abs(float(loss) - float(z['loss'])) < 1e-6 * float(z['loss']) + 1e-6
    loss.backward()
    _check_grads(m, {k: z['grad.' + k] for k in sd})


def test_tiny_golden_loss_through_torch_cross_entropy(golden_dir):
    """The unchanged trainer.py:426-434 loss code (torch CrossEntropyLoss) on our logits gives the same grads."""
    z = np.load(os.path.join(golden_dir, 'las_tiny.npz'))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd.')}
    x, lens, y = torch.from_numpy(z['x']), [int(v) for v in z['lens']], torch.from_numpy(z['y']).to(DEV)
    m = _model(tuple(int(v) for v in z['dims']), sd)
    ans_len = z['logits'].shape[1]
    _, pred, _ = m(x.to(DEV), ans_len, teacher=y, state_len=lens)
    label = y[:, 1:ans_len + 1].contiguous()
    b, t, c = pred.shape
    loss = torch.nn.CrossEntropyLoss(ignore_index=0, reduction='none')(pred.view(b * t, c), label.view(-1))
    loss = torch.mean(torch.sum(loss.view(b, t), dim=-1) / torch.sum(y != 0, dim=-1).to(dtype=torch.float32))
    loss.backward()
    assert abs(float(loss) - float(z['loss'])) < 1e-5
    _check_grads(m, {k: z['grad.' + k] for k in sd})


def test_tiny_golden_greedy_and_decode(golden_dir):
    z = np.load(os.path.join(golden_dir, 'las_tiny.npz'))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd.')}
    x, lens = torch.from_numpy(z['x']), [int(v) for v in z['lens']]
    m = _model(tuple(int(v) for v in z['dims']), sd)
    U = z['greedy_logits'].shape[1]
    with torch.no_grad():
        _, gl, ga = m(x.to(DEV), U, state_len=lens)
    assert np.array_equal(gl.argmax(-1).cpu().numpy(), z['greedy_logits'].argmax(-1))
    assert np.abs(gl.cpu().numpy() - z['greedy_logits']).max() < 1e-4
    assert np.abs(ga.numpy() - z['greedy_att']).max() < 1e-5

    class Mapper:          # the two ASRDataset.Mapper methods ASR.decode uses (asr.py:167,171)
        def ind_to_char(self, i):
            return O.TOKENS[i]

        def char_to_ind(self, c):
            return O.TOKENS.index(c)
    lm = _CharLM()
    lm.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('lm.')})
    for i in range(len(z['decode_lm0'])):
        xi = x[i:i + 1, :lens[i]].to(DEV)
        assert m.decode(xi, [lens[i]], None, Mapper(), 0.0) == str(z['decode_lm0'][i])
        got, want = m.decode(xi, [lens[i]], lm, Mapper(), 0.5), str(z['decode_lm05'][i])
        ids, margin = O.decode_greedy(sd, x[i:i + 1, :lens[i]], [lens[i]], {k[3:]: torch.from_numpy(z[k]) for k in z.files
                                                                             if k.startswith('lm.')}, 0.5, return_margin=True)
        n = min(len(got), len(want), 20)
        assert got[:n] == want[:n] and (margin < 1e-4 or got == want)


def test_default_dims_golden(golden_dir):
    from ss_asr_b200.functional import asr_loss
    z = np.load(os.path.join(golden_dir, 'las_default.npz'))
    dims = (50, 256, 256, 128, 80)
    torch.manual_seed(1)
    from ss_asr_b200.asr import ASR
    m = ASR(*dims, 1.0).to(DEV)                                       # same init stream as the reference
    x, lens, y = O.synth_batch(int(z['B']), int(z['T']), int(z['F']), int(z['U']), seed=1234)
    U = z['logits'].shape[1]
    el, logits, att = m(x.to(DEV), U, teacher=y.to(DEV), state_len=lens)
    assert el == [int(v) for v in z['enc_len']]
    assert np.abs(logits.detach().cpu().numpy() - z['logits']).max() < 1e-5
    assert np.abs(att.numpy() - z['att']).max() < 1e-5
    loss = asr_loss(logits, y.to(DEV))
    assert abs(float(loss) - float(z['loss'])) < 1e-6 * float(z['loss']) + 1e-6
    loss.backward()
    gtot = float(z['gnorm_total'])
    for k, p in m.named_parameters():
        g = p.grad.cpu()
        assert abs(float(g.double().norm()) - float(z['gnorm.' + k])) <= 1e-4 * float(z['gnorm.' + k]) + 1e-6 * gtot, k
        assert np.abs(g.flatten()[:256].numpy() - z['ghead.' + k]).max() <= 1e-4 * np.abs(z['ghead.' + k]).max() + 1e-6 * gtot, k


@pytest.mark.parametrize('dims,B,T,U,lens', [
    ((50, 32, 48, 16, 20), 7, 64, 9, None),
    ((50, 16, 32, 8, 10), 6, 91, 7, [91, 90, 75, 50, 23, 8]),        # odd T at every layer, length-1 encoder rows
    ((50, 48, 16, 24, 40), 1, 40, 5, [40]),                           # batch of one
    ((50, 16, 16, 8, 12), 70, 24, 4, None),                           # more utterances than one 64-row tile
])
def test_oracle_parity_shapes(dims, B, T, U, lens):
    from ss_asr_b200.functional import asr_loss
    sd = O.make_state_dict(*dims, seed=3)
    if lens is None:
        x, lens, y = O.synth_batch(B, T, dims[4], U, seed=77)
    else:
        x, _, y = O.synth_batch(B, T, dims[4], U, seed=77)
        for i, l in enumerate(lens):
            x[i, l:] = 0
            x[i, :l] += 0.01
    loss_o, logits_o, att_o, enc_o, grads_o = O.train_step_grads(sd, x, lens, y)
    m = _model(dims, sd)
    ans_len = logits_o.shape[1]
    el, logits, att = m(x.to(DEV), ans_len, teacher=y.to(DEV), state_len=lens)
    assert float((logits.detach().cpu() - logits_o).abs().max()) < 1e-5
    assert float((att - att_o).abs().max()) < 1e-5
    loss = asr_loss(logits, y.to(DEV))
    assert abs(float(loss) - float(loss_o)) < 1e-6 * float(loss_o) + 1e-6
    loss.backward()
    _check_grads(m, grads_o)


def test_blstm4_batch_coupling_quirk():
    """layer 4 recurs across utterances; perturbing the last utterance changes the first."""
    dims = (50, 16, 16, 8, 12)
    sd = O.make_state_dict(*dims, seed=1)
    m = _model(dims, sd)
    x, lens, _ = O.synth_batch(4, 32, 12, 4, seed=5)
    with torch.no_grad():
        e1, _ = m.encoder(x.to(DEV), lens)
        x2 = x.clone()
        x2[3, :lens[3]] += 0.5
        e2, _ = m.encoder(x2.to(DEV), lens)
        o2, _ = O.listener(sd, x2, lens)
    assert float((e1[0] - e2[0]).abs().max()) > 1e-4
    assert float((e2.cpu() - o2).abs().max()) < 1e-5


def test_decode_default_margin_strings(golden_dir):
    """Greedy transcripts identical to the unmodified reference's ASR.decode (bs=1 semantics, batched here)."""
    z = np.load(os.path.join(golden_dir, 'decode_default.npz'))
    sd = O.make_state_dict(50, 256, 256, 128, 80, seed=1)
    sd['char_trans.weight'] = sd['char_trans.weight'] * 20.0
    m = _model((50, 256, 256, 128, 80), sd)
    Ts = [int(v) for v in z['Ts']]
    xs = [torch.randn(1, Ti, 80, generator=torch.Generator().manual_seed(7000 + i)) for i, Ti in enumerate(Ts)]
    order = sorted(range(len(Ts)), key=lambda i: -Ts[i])
    xb = torch.zeros(len(Ts), max(Ts), 80)
    for j, i in enumerate(order):
        xb[j, :Ts[i]] = xs[i][0]
    ids = m.decode_batch(xb.to(DEV), [Ts[i] for i in order])
    for j, i in enumerate(order):
        assert O.ids_to_str(ids[j]) == str(z['margin_lm00'][i]), i
    lm = _CharLM()
    lm.load_state_dict(O.make_charlm_state_dict(50, 128, seed=7))
    for prec in ('fp32', 'tf32x3'):            # the CharLM-rescored loop on both exact paths (two-stream greedy loop, mode 3)
        ids = m.decode_batch(xb.to(DEV), [Ts[i] for i in order], rnn_lm=lm, lm_weight=0.5, precision=prec)
        for j, i in enumerate(order):
            assert O.ids_to_str(ids[j]) == str(z['margin_lm05'][i]), (prec, i)


def test_teacher_forcing_draws_follow_python_rng():
    """tf_rate < 1: one random.random() per step, like asr.py:94; sampled steps replayed through the oracle."""
    import random
    dims = (50, 16, 16, 8, 12)
    sd = O.make_state_dict(*dims, seed=1)
    m = _model(dims, sd, tf=0.5)
    x, lens, y = O.synth_batch(4, 32, 12, 8, seed=5)
    random.seed(11)
    _, logits, att = m(x.to(DEV), 9, teacher=y.to(DEV), state_len=lens)
    after = random.random()
    random.seed(11)
    mask = [random.random() <= 0.5 for _ in range(9)]
    assert random.random() == after and not all(mask) and any(mask)
    toks = m.last_tokens.cpu().long()                               # input token of every step
    sampled = torch.zeros(4, 9, dtype=torch.long)
    sampled[:, :8] = toks[:, 1:]
    with torch.no_grad():
        _, lo, ao, _ = O.asr_forward(sd, x, lens, 9, teacher=y, tf_mask=mask, sampled=sampled)
    assert float((logits.detach().cpu() - lo).abs().max()) < 1e-5
    for t in range(8):
        if mask[t]:
            assert torch.equal(toks[:, t + 1], y[:, t + 1])


# ------------------------------------------------------------------------------------------------ bf16 / tcgen05 path
@pytest.mark.parametrize('M,N,K,ak,bk', [(128, 128, 64, 0, 0), (300, 200, 80, 0, 0), (1024, 2048, 1024, 0, 0),
                                         (128, 128, 512, 16, 0), (256, 128, 1000, 0, 8), (2048, 80, 4096, 0, 0), (50, 256, 333, 0, 0),
                                         (8242, 1576, 328, 0, 0), (16384 + 128, 1024, 520, 8, 16),    # many tiles per SM, ragged edges
                                         (20000, 768, 256, 0, 0), (40000 + 77, 320, 192, 0, 8)])       # persistent 128 x 256 kernel, partial N tile
def test_gemm_bf16_tcgen05(M, N, K, ak, bk):
    """tcgen05+TMA GEMM against an fp64 product of the same bf16-rounded operands (fp32 accumulation tolerance)."""
    from ss_asr_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + N + K)
    Kp = (K + max(ak, bk) + 7) // 8 * 8
    Ab = torch.randn(M, Kp, generator=g).to(DEV).to(torch.bfloat16)
    Bb = torch.randn(N, Kp, generator=g).to(DEV).to(torch.bfloat16)
    bias = torch.randn(N, generator=g).to(DEV)
    ref = Ab[:, ak:ak + K].double() @ Bb[:, bk:bk + K].double().t() + bias.double()
    C = torch.zeros(M, N, device=DEV)
    _lib.check(lib.ssasr_gemm_bf16_tc(M, N, K, Ab.data_ptr(), Kp, ak, Bb.data_ptr(), Kp, bk, C.data_ptr(), N, bias.data_ptr(), 0,
                                      _lib.stream()), 'gemm_tc')
    assert float((C.double() - ref).abs().max()) < 2e-5 * K ** 0.5 * 4
    C2 = C.clone()
    _lib.check(lib.ssasr_gemm_bf16_tc(M, N, K, Ab.data_ptr(), Kp, ak, Bb.data_ptr(), Kp, bk, C2.data_ptr(), N, None, 1,
                                      _lib.stream()), 'gemm_tc accumulate')
    assert float((C2.double() - (2 * ref - bias.double())).abs().max()) < 4e-5 * K ** 0.5 * 4


def test_cvt_bf16_transposed_shifted_masked():
    from ss_asr_b200 import _lib
    lib = _lib.load()
    R, Cc, T = 96, 40, 12                       # 8 utterances x 12 frames
    src = torch.randn(R, Cc, device=DEV)
    Rp = (R + 7) // 8 * 8
    dst = torch.zeros(Cc, Rp, device=DEV, dtype=torch.bfloat16)
    _lib.check(lib.ssasr_cvt_bf16_t(src.data_ptr(), Cc, dst.data_ptr(), Rp, R, Cc, 0, 0, 0, 0, _lib.stream()), 'cvt_t')
    assert torch.equal(dst[:, :R], src.t().to(torch.bfloat16))


# bf16 training-path tolerances: logits atol 1e-2, loss rel 1e-3, gradients rel-L2 3e-2 and cosine >= 0.999
@pytest.mark.parametrize('dims,B,T,U', [((50, 256, 256, 128, 80), 8, 128, 20), ((50, 64, 64, 32, 40), 140, 48, 6),
                                        ((50, 512, 256, 128, 80), 3, 160, 6),        # C5 model: 512-dim BLSTM (64-row tiles)
                                        ((50, 128, 32, 16, 24), 5, 91, 7), ((50, 32, 48, 16, 20), 7, 64, 9)])
def test_bf16_training_path(dims, B, T, U):
    """tcgen05 gate GEMMs (+ tensor-core recurrence when S % 64 == 0) against the fp32 CPU oracle."""
    from ss_asr_b200.functional import asr_loss
    sd = O.make_state_dict(*dims, seed=1)
    x, lens, y = O.synth_batch(B, T, dims[4], U, seed=1234)
    loss_o, logits_o, att_o, enc_o, grads_o = O.train_step_grads(sd, x, lens, y)
    m = _model(dims, sd)
    m.train_precision = 'bf16'
    m.train()
    el, logits, att = m(x.to(DEV), logits_o.shape[1], teacher=y.to(DEV), state_len=lens)
    assert float((logits.detach().cpu() - logits_o).abs().max()) < 1e-2
    assert float((att - att_o).abs().max()) < 1e-3
    loss = asr_loss(logits, y.to(DEV))
    assert abs(float(loss) - float(loss_o)) < 1e-3 * float(loss_o)
    loss.backward()
    gtot = float(torch.sqrt(sum(v.double().pow(2).sum() for v in grads_o.values())))
    for k, p in m.named_parameters():
        a, b = p.grad.cpu().double(), grads_o[k].double()
        assert float((a - b).norm()) <= 3e-2 * float(b.norm()) + 1e-4 * gtot, k
        if float(b.norm()) > 1e-3 * gtot:
            assert float((a * b).sum() / (a.norm() * b.norm())) >= 0.999, k
    # eval / no-grad calls stay on the exact fp32 path even when train_precision is bf16
    m.eval()
    with torch.no_grad():
        _, gl, _ = m(x.to(DEV), 4, state_len=lens)
        _, gl_o, _, _ = O.asr_forward(sd, x, lens, 4)
    assert float((gl.cpu() - gl_o).abs().max()) < 1e-5


def test_tc_recurrence_many_tiles_per_cta():
    """Encoder with more 128-row batch tiles than tile groups (each CTA loops over several tiles)."""
    dims = (50, 256, 32, 16, 16)
    sd = O.make_state_dict(*dims, seed=2)
    B, T = 600, 16
    x, lens, _ = O.synth_batch(B, T, 16, 4, seed=9)
    m = _model(dims, sd)
    with torch.no_grad():
        eo, _ = O.listener(sd, x, lens)
    m.encoder.set_precision('bf16')
    xd = x.to(DEV).requires_grad_(True)
    e, el = m.encoder(xd, lens)
    assert float((e.detach().cpu() - eo).abs().max()) < 2e-2
    e.sum().backward()
    xg = x.clone().requires_grad_(True)
    p = {k: v.clone() for k, v in sd.items()}
    O.listener(p, xg, lens)[0].sum().backward()
    a, b = xd.grad.cpu().double(), xg.grad.double()
    assert float((a - b).norm()) <= 3e-2 * float(b.norm())


def test_decode_tf32x3_strings_and_gemm(golden_dir):
    """Opt-in tensor-core exact path (tf32 x 3 split) for the decode encoder: GEMM within 1e-4 of fp64 at K=1024 and
    greedy transcripts identical to the unmodified reference on the margin variant."""
    from ss_asr_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    M, N, K = 512, 384, 1024
    A = torch.randn(M, K, generator=g).to(DEV)
    B = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    C = torch.zeros(M, N, device=DEV)
    Aw, Bw = torch.empty(2 * M * K, device=DEV), torch.empty(2 * N * K, device=DEV)
    _lib.check(lib.ssasr_gemm_tf32x3(M, N, K, A.data_ptr(), Aw.data_ptr(), K, B.data_ptr(), Bw.data_ptr(), K, C.data_ptr(), N, None, 0,
                                     _lib.stream()), 'tf32x3')
    assert float((C.double() - A.double() @ B.double().t()).abs().max()) < 1e-4
    z = np.load(os.path.join(golden_dir, 'decode_default.npz'))
    sd = O.make_state_dict(50, 256, 256, 128, 80, seed=1)
    sd['char_trans.weight'] = sd['char_trans.weight'] * 20.0
    m = _model((50, 256, 256, 128, 80), sd)
    Ts = [int(v) for v in z['Ts']]
    xs = [torch.randn(1, Ti, 80, generator=torch.Generator().manual_seed(7000 + i)) for i, Ti in enumerate(Ts)]
    order = sorted(range(len(Ts)), key=lambda i: -Ts[i])
    xb = torch.zeros(len(Ts), max(Ts), 80)
    for j, i in enumerate(order):
        xb[j, :Ts[i]] = xs[i][0]
    ids = m.decode_batch(xb.to(DEV), [Ts[i] for i in order], precision='tf32x3')
    for j, i in enumerate(order):
        assert O.ids_to_str(ids[j]) == str(z['margin_lm00'][i]), i


@pytest.mark.parametrize('S,B,T', [(128, 37, 72), (256, 70, 136), (256, 3, 40), (512, 21, 72)])
def test_exact_cluster_recurrence_matches_fp32_path(S, B, T):
    """Forward-only exact path of the decode Listener: the split-operand quad-cluster recurrence (`rec_q_fwd_kernel<R,false,true>`:
    W_hi in tensor memory, W_lo in shared memory, hi + lo h images exchanged inside the cluster) against the fp32 SIMT path
    (itself pinned to the reference goldens) and against the counter-barrier X3 kernels it replaces (S = 512: the 16-CTA cluster
    kernel `rec_wide_fwd_kernel<true>`, which replaces the SIMT recurrence there): ragged lengths, partial
    batch tiles, both tile heights, batched (seq-first `blstm_4` quirk, asr.py:262) and per-utterance `blstm_4`.
    Tolerance 5e-6 on |h| <= 1 (measured 4e-7 .. 9e-7, the same as the kernels it replaces)."""
    dims = (50, S, 32, 16, 24)
    sd = O.make_state_dict(*dims, seed=2)
    x, lens, _ = O.synth_batch(B, T, 24, 4, seed=5)
    m = _model(dims, sd).eval()
    xd = x.to(DEV)

    def enc(prec, env):
        if env is None:
            os.environ.pop('SSASR_REC_Q_X3', None)
        else:
            os.environ['SSASR_REC_Q_X3'] = env
        m.encoder.set_precision(prec)
        try:
            with torch.no_grad():
                e, el = m.encoder(xd, lens)
            torch.cuda.synchronize()
        finally:
            m.encoder.set_precision('fp32')
            os.environ.pop('SSASR_REC_Q_X3', None)
        return e

    for indep in (False, True):
        m.encoder.utterance_independent = indep
        ref = enc('fp32', None)
        old = enc('tf32x3', '0')
        for env in (None, '16', '32'):
            new = enc('tf32x3', env)
            assert not bool(torch.isnan(new).any())
            assert float((new - ref).abs().max()) < 5e-6, (indep, env)
            assert float((new - old).abs().max()) < 2e-6, (indep, env)
    m.encoder.utterance_independent = False


def test_per_step_module_api_matches_fused_forward():
    """Attention.forward / Speller.forward called one step at a time (text_autoencoder.py:52-94 usage) reproduce the
    oracle's logits and gradients, i.e. the TAE/SAE/ADV trainers' call pattern keeps working on the drop-in modules."""
    dims = (50, 32, 48, 16, 20)
    sd = O.make_state_dict(*dims, seed=3)
    x, lens, y = O.synth_batch(5, 64, 20, 6, seed=21)
    loss_o, logits_o, att_o, enc_o, grads_o = O.train_step_grads(sd, x, lens, y)
    m = _model(dims, sd)
    U = logits_o.shape[1]
    yd = y.to(DEV)
    enc, enc_len = m.encoder(x.to(DEV), lens)
    teacher = m.embed(yd)
    m.decoder.init_rnn(enc.shape[0], enc.device)
    m.attention.reset_enc_mem()
    last = m.embed(torch.zeros(enc.shape[0], dtype=torch.long, device=DEV))
    outs, atts = [], []
    for t in range(U):
        a, ctxv = m.attention(m.decoder.state_list[0], enc, enc_len)
        dec_out = m.decoder(torch.cat([last, ctxv], dim=-1))
        outs.append(m.char_trans(dec_out))
        atts.append(a)
        last = teacher[:, t + 1, :]
    logits = torch.stack(outs, 1)
    assert float((logits.detach().cpu() - logits_o).abs().max()) < 1e-5
    assert float((torch.stack(atts, 1).detach().cpu() - att_o).abs().max()) < 1e-5
    assert m.attention.comp_listener_feature is not None and m.attention.state_mask.dtype == torch.bool
    h, c = m.decoder.hidden_state
    assert len(h) == 2 and not h[0].is_cuda
    from ss_asr_b200.functional import asr_loss
    asr_loss(logits, yd).backward()
    _check_grads(m, grads_o)


@pytest.mark.parametrize('S,B,T,K', [(256, 200, 48, 64), (256, 37, 20, 1024), (128, 70, 33, 40), (64, 9, 21, 24)])
def test_cluster_recurrence_matches_counter_barrier_kernels(S, B, T, K):
    """The cluster recurrent kernels of rec_cl.cu -- the quad formulation (64 units per CTA, transposed product, 16- and 32-row
    tiles; S in {128, 256}) and the 8-CTA / 64-row kernels -- against the counter-barrier kernels (rec_tc.cu) on ragged
    lengths, with (K <= 80) and without the fused input projection: same bf16 operand rounding and fp32 accumulation, so they
    agree far inside the bf16-path tolerance; the debug stamps prove which kernel ran."""
    from ss_asr_b200 import _lib
    from ss_asr_b200.asr import pBLSTM
    lib = _lib.load()
    assert lib.ssasr_rec_cl_capacity(S, 0) >= 2 * ((B + 63) // 64) and lib.ssasr_rec_cl_capacity(S, 1) >= 2 * ((B + 63) // 64)
    torch.manual_seed(3)
    m = pBLSTM(K, S).to(DEV)
    m.precision = 'bf16'
    g = torch.Generator().manual_seed(5)
    lens = sorted([int(v) for v in torch.randint(max(2, T // 3), T + 1, (B,), generator=g)], reverse=True)
    lens[0] = T
    x0 = torch.randn(B, T, K, generator=g)
    x0 = x0 * (torch.arange(T)[None, :, None] < torch.tensor(lens)[:, None, None])
    res = {}
    stamps = torch.zeros(T + 2, 12, dtype=torch.int64, device=DEV)     # one row per step; odd T runs T + 1 (padded) steps
    try:
        for name, rows, on in (('quad16', 16, 1), ('quad16_ring', 16, 1), ('quad32', 32, 1), ('cl8', 0, 1), ('counter', 0, 0)):
            lib.ssasr_rec_q_set_rows(rows)
            lib.ssasr_rec_cl_enable(on)
            lib.ssasr_rec_set_dsmem(0 if name.endswith('_ring') else 1)      # cluster exchange through the L2 ring / by DSMEM copies
            stamps.zero_()
            lib.ssasr_rec_cl_set_debug(stamps.data_ptr())
            x = x0.clone().to(DEV).requires_grad_(True)
            m.zero_grad()
            out, _, _ = m(x, state_len=lens, pack_input=True)
            w = torch.linspace(-1, 1, out.numel(), device=DEV).view_as(out)
            (out * w).sum().backward()
            torch.cuda.synchronize()
            lib.ssasr_rec_cl_set_debug(None)
            ran_cluster = bool((stamps != 0).any())
            assert ran_cluster == bool(on)
            res[name] = (out.detach().clone(), x.grad.clone(), [p.grad.clone() for p in m.parameters()])
    finally:
        lib.ssasr_rec_q_set_rows(16)
        lib.ssasr_rec_cl_enable(1)
        lib.ssasr_rec_set_dsmem(1)
        lib.ssasr_rec_cl_set_debug(None)
    b = res['counter']
    for name in ('quad16', 'quad16_ring', 'quad32', 'cl8'):
        a = res[name]
        assert float((a[0] - b[0]).abs().max()) < 2e-3, name
        assert float(a[1].norm()) > 0 and float((a[1] - b[1]).norm()) <= 5e-3 * float(b[1].norm()), name
        for ga, gb in zip(a[2], b[2]):
            assert float((ga - gb).norm()) <= 5e-3 * float(gb.norm()) + 1e-6, name


def test_host_batch_pipeline_and_async_attention_maps(golden_dir):
    """The e2e helpers: double-buffered pinned-host -> device batches and the non-blocking D2H of the attention maps
    deliver the same values as the blocking reference behaviour (asr.py:104 `.cpu()`)."""
    from ss_asr_b200.parallel import HostBatchPipeline
    dims = (50, 64, 64, 32, 40)
    sd = O.make_state_dict(*dims, seed=1)
    x, lens, y = O.synth_batch(6, 48, dims[4], 7, seed=11)
    m = _model(dims, sd)
    m.eval()
    pipe = HostBatchPipeline(DEV)
    xh, yh = x.pin_memory(), y.pin_memory()
    pipe.submit(xh, yh)
    pipe.submit(xh, yh)
    outs = []
    with torch.no_grad():
        for use_async in (False, True):
            xd, yd = pipe.take()
            assert xd.is_cuda and torch.equal(xd.cpu(), x) and torch.equal(yd.cpu(), y)
            m.att_async = use_async
            _, logits, att = m(xd, 8, teacher=yd, state_len=lens)
            torch.cuda.synchronize()
            assert not att.is_cuda
            outs.append((logits.cpu().clone(), att.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_fused_adadelta_clip_step_matches_torch():
    """Solver.step (trainer.py:131-148) fused on the device against clip_grad_norm_ + torch.optim.Adadelta, several steps,
    odd tensor sizes, clipping active and inactive, and the NaN-skip branch."""
    from ss_asr_b200.optim import FusedAdadelta
    g = torch.Generator().manual_seed(0)
    shapes = [(1,), (3,), (257, 129), (70001,), (64, 1024), (5, 7, 3)]
    ref_p = [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ref = torch.optim.Adadelta(ref_p, lr=1.0, eps=1e-8)
    ours = FusedAdadelta(our_p, lr=1.0, eps=1e-8)
    for it, scale in enumerate([10.0, 0.01, 3.0, 1.0]):          # norm far above / far below the clip threshold
        grads = [scale * torch.randn(*s, generator=g).to(DEV) for s in shapes]
        for p, q, gr in zip(ref_p, our_p, grads):
            p.grad = gr.clone()
            q.grad = gr.clone()
        gn = torch.nn.utils.clip_grad_norm_(ref_p, 5.0)
        ref.step()
        got = ours.step_clipped(5.0, write_clipped_grads=(it == 0))
        assert abs(float(got) - float(gn)) <= 1e-5 * float(gn)
        assert float(ours.last_applied) == 1.0
        for p, q in zip(ref_p, our_p):
            assert float((p - q).abs().max()) <= 2e-6 * (1.0 + float(p.abs().max()))
        if it == 0:
            for p, q in zip(ref_p, our_p):
                assert float((p.grad - q.grad).abs().max()) <= 1e-6 * (1.0 + float(p.grad.abs().max()))
    for k in ('square_avg', 'acc_delta'):
        for p, q in zip(ref_p, our_p):
            a, b = ref.state[p][k], ours.state[q][k]
            assert float((a - b).abs().max()) <= 1e-5 * (1e-6 + float(a.abs().max()))
    # NaN gradient: the step is cancelled on the device
    before = [q.detach().clone() for q in our_p]
    for q in our_p:
        q.grad = torch.randn_like(q)
    our_p[2].grad[0, 0] = float('nan')
    ours.step_clipped(5.0)
    assert float(ours.last_applied) == 0.0 and bool(torch.isnan(ours.last_grad_norm))
    for b, q in zip(before, our_p):
        assert torch.equal(b, q.detach())
    # state_dict round trip into torch.optim.Adadelta
    ref2 = torch.optim.Adadelta(our_p, lr=1.0, eps=1e-8)
    ref2.load_state_dict(ours.state_dict())


def test_deferred_weight_gradients_match():
    """functional.set_overlap_wgrad: Speller and encoder weight-gradient GEMMs on a second stream, joined by the consumers
    (FusedAdadelta.step_clipped / join_deferred).  Same gradients and same parameter update as the in-order schedule."""
    from ss_asr_b200 import functional as Fk
    from ss_asr_b200.functional import asr_loss
    from ss_asr_b200.optim import FusedAdadelta
    dims = (50, 64, 64, 32, 40)
    sd = O.make_state_dict(*dims, seed=1)
    x, lens, y = O.synth_batch(70, 64, dims[4], 6, seed=21)
    out = {}
    try:
        for on in (False, True):
            Fk.set_overlap_wgrad(on)
            m = _model(dims, sd)
            m.train_precision = 'bf16'
            m.train()
            opt = FusedAdadelta(m.parameters(), lr=1.0, eps=1e-8)
            for _ in range(2):
                opt.zero_grad(set_to_none=True)
                _, logits, _ = m(x.to(DEV), 7, teacher=y.to(DEV), state_len=lens)
                n0 = Fk._OVERLAP['deferred_total']
                asr_loss(logits, y.to(DEV)).backward()
                if on:
                    assert Fk._OVERLAP['deferred_total'] - n0 == 5   # the Speller's + one deferred batch per encoder layer
                # joined when the autograd engine finished the backward pass: any reader of .grad is ordered after the side stream
                assert len(Fk._OVERLAP['pending']) == 0
                grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
                opt.step_clipped(5.0)
            torch.cuda.synchronize()
            out[on] = (grads, {k: p.detach().clone() for k, p in m.named_parameters()})
    finally:
        Fk.set_overlap_wgrad(False)
    for k in out[False][0]:
        a, b = out[True][0][k], out[False][0][k]
        assert float((a - b).norm()) <= 1e-4 * float(b.norm()) + 1e-7, k
        a, b = out[True][1][k], out[False][1][k]
        assert float((a - b).abs().max()) <= 1e-4 * (1e-3 + float(b.abs().max())), k


def test_prepare_x_prepare_y_match_reference_glue():
    """ASRDataset.py:297-340 restated inline (numpy count on the host) against the device-side prepare_x / prepare_y."""
    from ss_asr_b200.dataset import prepare_x, prepare_y
    g = torch.Generator().manual_seed(7)
    B, T, F = 9, 57, 40
    lens = sorted([int(v) for v in torch.randint(5, T + 1, (B,), generator=g)], reverse=True)
    x = torch.randn(1, B, T, F, generator=g, dtype=torch.float64) * 4 - 10
    x = x * (torch.arange(T)[None, None, :, None] < torch.tensor(lens)[None, :, None, None])
    y = torch.randint(0, 30, (1, B, 13), generator=g).to(torch.float64)
    for dt in (torch.float64, torch.float32):
        xr = x.to(dt).squeeze(0).to(torch.float32)
        want_lens = [int(v) for v in np.sum(np.sum(xr.numpy(), axis=-1) != 0, axis=-1)]
        got, got_lens = prepare_x(x.to(dt), device=torch.device(DEV))
        assert got.dtype == torch.float32 and got.is_cuda and torch.equal(got.cpu(), xr)
        assert got_lens == want_lens == lens
    yr = y.squeeze(0).to(torch.long)
    gy, gl = prepare_y(y, device=torch.device(DEV))
    assert torch.equal(gy.cpu(), yr) and gl == [int(v) + 1 for v in torch.sum(yr != 0, dim=-1)]


# ------------------------------------------------------------------------------------------------ validation metrics (f4)
class _Mapper:
    """ASRDataset.Mapper (ASRDataset.py:228-262) over the default token table."""

    def __init__(self, tokens=O.TOKENS):
        self.mapping = {c: i for i, c in enumerate(tokens)}

    def char_to_ind(self, c):
        return self.mapping[c]

    def ind_to_char(self, i):
        return O.TOKENS[i]


def test_calc_acc_err_match_oracle_and_reference_golden(golden_dir):
    """Integer work: per-utterance counts bit-exact against the oracle; batch results identical (==) to the committed outputs
    of the unmodified reference."""
    from oracle import postprocess_oracle as PO
    from ss_asr_b200 import postprocess as PPc
    z = np.load(os.path.join(golden_dir, 'postprocess.npz'))
    for k in range(4):
        seed, B, U, L = (int(v) for v in z['case_%d' % k])
        predict, label, pred_tok = PO.synth_cases(seed=seed, B=B, U=U, L=L)
        pd_, ld_ = torch.from_numpy(predict).to(DEV), torch.from_numpy(label).to(DEV)
        stats, toks = PPc.utterance_stats(pd_, ld_, _Mapper(), return_tokens=True)
        assert np.array_equal(toks.cpu().numpy(), pred_tok)
        assert np.array_equal(stats.numpy(), PO.utterance_stats(predict, label))
        assert PPc.calc_acc(pd_, ld_) == float(z['acc_%d' % k])
        assert PPc.calc_err(pd_, ld_, _Mapper()) == float(z['err_%d' % k])
        assert PPc.calc_acc_err(pd_, ld_, _Mapper()) == (float(z['acc_%d' % k]), float(z['err_%d' % k]))


def test_calc_acc_err_edge_cases():
    from oracle import postprocess_oracle as PO
    from ss_asr_b200 import postprocess as PPc
    # NaN rows (np.argmax: the first NaN wins), -inf rows (index 0), long sequences, strided (sliced) inputs
    rng = np.random.RandomState(5)
    predict, label, _ = PO.synth_cases(seed=21, B=7, U=300, L=260)
    predict[0, 3, 17] = np.nan
    predict[0, 3, 40] = np.nan
    predict[1, 0, :] = -np.inf
    predict[2, 5, :] = 1.25
    want = PO.utterance_stats(predict, label)
    big = torch.from_numpy(rng.randn(7, 310, 64).astype(np.float32)).to(DEV)
    big[:, :300, :50] = torch.from_numpy(predict).to(DEV)
    lab = torch.zeros(7, 270, dtype=torch.int64, device=DEV)
    lab[:, 5:265] = torch.from_numpy(label).to(DEV)
    stats, toks = PPc.utterance_stats(big[:, :300, :50], lab[:, 5:265], None, return_tokens=True)
    assert np.array_equal(toks.cpu().numpy(), np.argmax(predict, axis=-1))
    assert np.array_equal(stats.numpy(), want)
    # an empty label (first token 0) divides by zero in the reference (postprocess.py:27); same exception here
    lab0 = torch.zeros(2, 4, dtype=torch.int64, device=DEV)
    with pytest.raises(ZeroDivisionError):
        PPc.calc_acc(torch.zeros(2, 4, 50, device=DEV), lab0)
    with pytest.raises(RuntimeError):
        PPc.calc_acc(torch.zeros(2, 4, 50), lab0.cpu())


def test_valid_step_matches_oracle():
    """Body of ASRTrainer.valid (trainer.py:472-494): batched greedy forward for ans_len + 30 steps without a teacher, loss
    on the first ans_len steps, calc_acc / calc_err -- against the oracle's no-teacher forward."""
    from oracle import postprocess_oracle as PO
    from ss_asr_b200 import postprocess as PPc
    dims = (50, 32, 32, 16, 20)
    sd = O.make_state_dict(*dims, seed=1)
    sd['char_trans.weight'] = sd['char_trans.weight'] * 20          # margins between classes (SURVEY §8d "margin" variant)
    x, lens, y = O.synth_batch(6, 48, 20, 6, seed=3)
    y_lens = [int(v) + 1 for v in (y != 0).sum(-1)]
    ans_len = max(y_lens) - 1
    _, logits_o, att_o, _ = O.asr_forward(sd, x, lens, ans_len + 30)
    loss_o = O.asr_loss(logits_o[:, :ans_len], y)
    label = y[:, 1:ans_len + 1]
    m = _model(dims, sd).eval()
    loss, acc, err, pred, att, toks = PPc.valid_step(m, x.to(DEV), y.to(DEV), lens, y_lens, _Mapper())
    assert float((pred.cpu() - logits_o).abs().max()) < 1e-4
    assert abs(float(loss) - float(loss_o)) < 1e-5 * abs(float(loss_o)) + 1e-6
    assert np.array_equal(toks.cpu().numpy(), np.argmax(logits_o.numpy(), -1))
    assert acc == PO.calc_acc(logits_o.numpy(), label.numpy())
    assert err == PO.calc_err(logits_o.numpy(), label.numpy())


def test_decode_batch_encoder_chunking_is_invisible():
    """decode_batch runs the Listener over length-sorted groups of utterances (each only as deep as its longest member);
    tokens must be identical to the single-pass result and to the oracle's bs=1 decode, on both exact paths."""
    dims = (50, 64, 64, 32, 40)
    sd = O.make_state_dict(*dims, seed=1)
    sd['char_trans.weight'] = sd['char_trans.weight'] * 20
    g = torch.Generator().manual_seed(7)
    Ts = sorted([int(v) for v in torch.randint(17, 72, (11,), generator=g)], reverse=True)
    xb = torch.zeros(len(Ts), Ts[0], 40)
    for i, t in enumerate(Ts):
        xb[i, :t] = torch.randn(t, 40, generator=g)
    m = _model(dims, sd).eval()
    want = [O.decode_greedy(sd, xb[i:i + 1, :t], [t], max_steps=12) for i, t in enumerate(Ts)]
    for prec in ('fp32', 'tf32x3'):
        outs = []
        for chunk in (0, 4, 5, 64):
            m.decode_encoder_chunk = chunk
            outs.append(m.decode_batch(xb.to(DEV), Ts, max_steps=12, precision=prec))
        assert outs[0] == outs[1] == outs[2] == outs[3], prec
        assert outs[0] == [list(w) for w in want], prec
        # a HOST batch (pinned or pageable) goes through the pipelined per-group upload (strided copy of the group's frames)
        for chunk in (0, 4):
            m.decode_encoder_chunk = chunk
            assert m.decode_batch(xb.pin_memory(), Ts, max_steps=12, precision=prec) == outs[0], (prec, chunk)
        assert m.decode_batch(xb.clone(), Ts, max_steps=12, precision=prec) == outs[0], prec


def test_beam_search_matches_oracle_and_reduces_to_greedy(golden_dir):
    """beam search (csrc/beam.cu + ASR.beam_decode_batch).  The reference configures a beam but decodes greedily
    (trainer.py:590), so the anchor on the reference is beam_size 1 == its greedy strings (decode_default.npz, margin variant, with
    and without the CharLM); for beam sizes 3 and 8 the CUDA path must return the oracle's (`O.decode_beam`) best hypothesis and
    score (1e-3: sums of up to 25 fp32 log-probabilities), with ragged lengths and the LM."""
    z = np.load(os.path.join(golden_dir, 'decode_default.npz'))
    sd = O.make_state_dict(50, 256, 256, 128, 80, seed=1)
    sd['char_trans.weight'] = sd['char_trans.weight'] * 20.0
    m = _model((50, 256, 256, 128, 80), sd)
    Ts = [int(v) for v in z['Ts']]
    xs = [torch.randn(1, Ti, 80, generator=torch.Generator().manual_seed(7000 + i)) for i, Ti in enumerate(Ts)]
    order = sorted(range(len(Ts)), key=lambda i: -Ts[i])
    xb = torch.zeros(len(Ts), max(Ts), 80)
    for j, i in enumerate(order):
        xb[j, :Ts[i]] = xs[i][0]
    lens = [Ts[i] for i in order]
    ids = m.beam_decode_batch(xb.to(DEV), lens, 1)
    for j, i in enumerate(order):
        assert O.ids_to_str(ids[j]) == str(z['margin_lm00'][i]), i
    assert ids == m.decode_batch(xb.to(DEV), lens)
    # small model, beams 3 and 8, with and without the LM, against the oracle
    dims = (50, 64, 64, 32, 40)
    sd = O.make_state_dict(*dims, seed=5)
    sd['char_trans.weight'] = sd['char_trans.weight'] * 8.0
    lm_sd = O.make_charlm_state_dict(50, 128, seed=7)
    lm = _charlm_module(lm_sd)
    g = torch.Generator().manual_seed(3)
    Ts = sorted([int(v) for v in torch.randint(17, 70, (7,), generator=g)], reverse=True)
    xb = torch.zeros(len(Ts), Ts[0], 40)
    for i, t in enumerate(Ts):
        xb[i, :t] = torch.randn(t, 40, generator=g)
    m = _model(dims, sd).eval()
    for W in (3, 8):
        for use_lm in (False, True):
            got, sc = m.beam_decode_batch(xb.to(DEV), Ts, W, max_steps=25, rnn_lm=lm if use_lm else None,
                                          lm_weight=0.5 if use_lm else 0.0, return_scores=True)
            for i, t in enumerate(Ts):
                want, ws = O.decode_beam(sd, xb[i:i + 1, :t], [t], W, lm=lm_sd if use_lm else None,
                                         lm_weight=0.5 if use_lm else 0.0, max_steps=25, return_score=True)
                assert got[i] == list(want), (W, use_lm, i)
                assert abs(sc[i] - ws) < 1e-3 * max(1.0, abs(ws)), (W, use_lm, i, sc[i], ws)
    # the single-utterance entry point with the reference's Mapper-style object
    class _Map:
        def char_to_ind(self, c):
            return O.TOKENS.index(c)

        def ind_to_char(self, i):
            return O.TOKENS[i]
    s3 = m.beam_decode(xb[:1].to(DEV), Ts[:1], lm, _Map(), 0.5, 3)
    assert s3 == O.ids_to_str(O.decode_beam(sd, xb[:1, :Ts[0]], Ts[:1], 3, lm=lm_sd, lm_weight=0.5))


def test_greedy_loop_two_streams_is_invisible():
    """Greedy decoding runs the layer-2 chain of step t (gate product, cell, character projection + argmax, embedding row of the
    next input) on a second stream under the attention step t+1 (asr.py:84: the query is the LAYER-1 state): tokens must be
    identical to the single-stream loop (SSASR_DECODE_DUAL=0) and to the oracle's bs=1 decode, on both exact paths, with the
    EOS check at several periods, several times in a row (race check)."""
    dims = (50, 64, 64, 32, 40)
    sd = O.make_state_dict(*dims, seed=3)
    sd['char_trans.weight'] = sd['char_trans.weight'] * 20
    g = torch.Generator().manual_seed(11)
    Ts = sorted([int(v) for v in torch.randint(17, 90, (37,), generator=g)], reverse=True)
    xb = torch.zeros(len(Ts), Ts[0], 40)
    for i, t in enumerate(Ts):
        xb[i, :t] = torch.randn(t, 40, generator=g)
    m = _model(dims, sd).eval()
    want = [list(O.decode_greedy(sd, xb[i:i + 1, :t], [t], max_steps=25)) for i, t in enumerate(Ts[:6])]
    try:
        for prec in ('fp32', 'tf32x3'):
            os.environ['SSASR_DECODE_DUAL'] = '0'
            single = m.decode_batch(xb.to(DEV), Ts, max_steps=25, precision=prec)
            os.environ.pop('SSASR_DECODE_DUAL')
            assert single[:6] == want, prec
            for check in (16, 1, 0, 7):
                m.decode_stop_check = check
                for _ in range(3):
                    assert m.decode_batch(xb.to(DEV), Ts, max_steps=25, precision=prec) == single, (prec, check)
        # with the CharLM: its recurrences of step t run on a THIRD stream (they need token t only), the layer-2 stream mixes
        # and selects; against one stream and against the fused LM kernel on the layer-2 stream
        lm = _charlm_module(O.make_charlm_state_dict(50, 128, seed=7))
        m.decode_stop_check = 16
        for prec in ('fp32', 'tf32x3'):
            os.environ['SSASR_DECODE_DUAL'] = '0'
            single = m.decode_batch(xb.to(DEV), Ts, max_steps=25, rnn_lm=lm, lm_weight=0.5, precision=prec)
            os.environ.pop('SSASR_DECODE_DUAL')
            os.environ['SSASR_DECODE_LM_SPLIT'] = '0'
            assert m.decode_batch(xb.to(DEV), Ts, max_steps=25, rnn_lm=lm, lm_weight=0.5, precision=prec) == single, prec
            os.environ.pop('SSASR_DECODE_LM_SPLIT')
            for _ in range(3):
                assert m.decode_batch(xb.to(DEV), Ts, max_steps=25, rnn_lm=lm, lm_weight=0.5, precision=prec) == single, prec
    finally:
        os.environ.pop('SSASR_DECODE_DUAL', None)
        os.environ.pop('SSASR_DECODE_LM_SPLIT', None)
        m.decode_stop_check = 16


@pytest.mark.parametrize('tf_rate', [1.0, 0.6])
def test_dual_stream_speller_is_invisible(tf_rate):
    """bf16 path: the Speller's layer-2 chain on the internal second stream (forward and backward) runs the same kernels on the
    same data as the single-stream loop -- logits, sampled tokens and attention maps must be bit-identical, gradients equal up to
    the summation order of the split-K weight-gradient GEMMs,
    with teacher forcing and with sampled steps (which join the two streams), several times in a row (race check)."""
    import random
    from ss_asr_b200 import functional as Fk
    from ss_asr_b200.functional import asr_loss
    dims = (50, 64, 64, 32, 40)
    sd = O.make_state_dict(*dims, seed=1)
    x, lens, y = O.synth_batch(70, 48, 40, 17, seed=11)

    def run(dual):
        Fk.set_dual_stream_speller(dual)
        m = _model(dims, sd, tf=tf_rate)
        m.train_precision = 'bf16'
        m.train()
        m.sample_seed = 5
        random.seed(3)
        _, logits, att = m(x.to(DEV), 18, teacher=y.to(DEV), state_len=lens)
        asr_loss(logits, y.to(DEV)).backward()
        torch.cuda.synchronize()
        return logits.detach().clone(), att.clone(), m.last_tokens.clone(), {k: p.grad.clone() for k, p in m.named_parameters()}

    try:
        want = run(False)
        for _ in range(3):
            got = run(True)
            assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]) and torch.equal(got[2], want[2])
            for k in want[3]:       # weight gradients go through split-K atomics: run-to-run identical only up to fp32 summation order
                a, b = got[3][k].double(), want[3][k].double()
                assert float((a - b).norm()) <= 1e-5 * float(b.norm()) + 1e-12, k
    finally:
        Fk.set_dual_stream_speller(True)


def test_error_behaviour_matches_reference_contract():
    """SURVEY §8b 'Errors': unsorted / zero lengths and T < 8 raise RuntimeError (pack_padded_sequence / nn.LSTM in the
    reference, asr.py:413-418), missing lengths and bs != 1 decoding assert (asr.py:411, asr.py:125)."""
    dims = (50, 16, 16, 8, 12)
    sd = O.make_state_dict(*dims, seed=1)
    m = _model(dims, sd)
    x, lens, y = O.synth_batch(3, 32, 12, 4, seed=5)
    with pytest.raises(RuntimeError):
        m(x.to(DEV), 4, teacher=y.to(DEV), state_len=[20, 32, 25])          # not sorted in decreasing order
    with pytest.raises(RuntimeError):
        m(x.to(DEV), 4, teacher=y.to(DEV), state_len=[32, 20, 0])           # zero-length utterance
    with pytest.raises(RuntimeError):
        m(x[:, :7].to(DEV), 4, teacher=y.to(DEV), state_len=[7, 7, 7])      # fewer than 8 frames: nothing left after 3 halvings
    with pytest.raises(AssertionError):
        m.encoder.blstm_1(x.to(DEV), state_len=None, pack_input=True)
    with pytest.raises(AssertionError):
        m.decode(x.to(DEV), lens, None, None, 0.0)                          # ASR.decode is bs=1 only
    # the shortest usable input: 8 frames -> one encoder frame
    x8, l8, y8 = O.synth_batch(2, 8, 12, 3, seed=6)
    l8 = [8, 8]
    _, logits_o, att_o, _ = O.asr_forward(sd, x8, l8, 3, teacher=y8)
    _, logits, att = m(x8.to(DEV), 3, teacher=y8.to(DEV), state_len=l8)
    assert float((logits.detach().cpu() - logits_o).abs().max()) < 1e-5 and att.shape[-1] == 1


def test_decode_batch_stops_when_every_utterance_has_emitted_eos():
    """decode_batch leaves the attend-and-spell loop once all utterances have produced EOS (asr.py:161-162 per utterance);
    the transcripts are those of the full-length loop and of the oracle's bs=1 decode."""
    dims = (50, 32, 32, 16, 20)
    sd = O.make_state_dict(*dims, seed=1)
    sd['char_trans.weight'] = sd['char_trans.weight'] * 20
    g = torch.Generator().manual_seed(9)
    Ts = sorted([int(v) for v in torch.randint(16, 60, (9,), generator=g)], reverse=True)
    xb = torch.zeros(len(Ts), Ts[0], 20)
    for i, t in enumerate(Ts):
        xb[i, :t] = torch.randn(t, 20, generator=g)
    for eos_bias, expect_early in ((0.0, False), (2.5, True), (50.0, True)):
        sd2 = dict(sd)
        b = sd['char_trans.bias'].clone()
        b[1] = eos_bias
        sd2['char_trans.bias'] = b
        m = _model(dims, sd2).eval()
        want = [O.decode_greedy(sd2, xb[i:i + 1, :t], [t], max_steps=40) for i, t in enumerate(Ts)]
        m.decode_stop_check = 0
        full = m.decode_batch(xb.to(DEV), Ts, max_steps=40)
        assert m.last_decode_steps == 41
        m.decode_stop_check = 4
        early = m.decode_batch(xb.to(DEV), Ts, max_steps=40)
        assert full == early == [list(w) for w in want], eos_bias
        if expect_early and max(len(w) for w in want) < 30:
            assert m.last_decode_steps % 4 == 0 and max(len(w) for w in want) < m.last_decode_steps <= max(len(w) for w in want) + 5
        if eos_bias == 50.0:
            assert all(len(w) == 0 for w in want) and m.last_decode_steps == 4


def test_cluster_recurrence_on_many_streams():
    """The exchange rings of the cluster recurrent kernels are pooled per (device, stream): more distinct streams than pool
    entries (16) must keep working (the oldest entry is handed over) and give the same result as the default stream."""
    dims = (50, 128, 32, 16, 24)
    sd = O.make_state_dict(*dims, seed=2)
    x, lens, _ = O.synth_batch(9, 40, 24, 4, seed=5)
    m = _model(dims, sd).eval()
    xd = x.to(DEV)
    m.encoder.set_precision('tf32x3')
    try:
        with torch.no_grad():
            ref, _ = m.encoder(xd, lens)
            torch.cuda.synchronize()
            for _ in range(20):
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    e, _ = m.encoder(xd, lens)
                s.synchronize()
                assert torch.equal(e, ref)
    finally:
        m.encoder.set_precision('fp32')
