"""Generates the committed golden vectors by running the UNMODIFIED reference
(/root/reference/src/asr.py, charlm.py through ref_shim) on seeded synthetic inputs.

Run in the authoring container only:   python tests/golden/make_golden.py
Outputs: tests/golden/las_tiny.npz, las_default.npz, decode_default.npz, fbank_1s.npz
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_shim  # noqa: E402
from oracle import las_oracle as O  # noqa: E402
from oracle import fbank_oracle as FB  # noqa: E402

torch.set_num_threads(8)


class _Mapper:  # ASRDataset.Mapper semantics (ASRDataset.py:228-262) without pandas import chain
    def char_to_ind(self, c):
        return O.TOKENS.index(c)

    def ind_to_char(self, i):
        return O.TOKENS[i]


def ref_model(asr_mod, dims, sd=None, seed=1):
    torch.manual_seed(seed)
    m = asr_mod.ASR(*dims)
    if sd is not None:
        m.load_state_dict(sd)
    return m


def run_train(asr_mod, dims, sd, x, lens, y):
    m = ref_model(asr_mod, dims, sd)
    y_lens = [int(l) + 1 for l in torch.sum(y != 0, dim=-1)]
    ans_len = max(y_lens) - 1
    random.seed(1)
    el, pred, att = m(x, ans_len, teacher=y, state_len=lens)
    label = y[:, 1:ans_len + 1].contiguous()
    b, t, c = pred.shape
    lossf = torch.nn.CrossEntropyLoss(ignore_index=0, reduction='none')
    loss = lossf(pred.view(b * t, c), label.view(-1))
    loss = torch.sum(loss.view(b, t), dim=-1) / torch.sum(y != 0, dim=-1).to(dtype=torch.float32)
    loss = torch.mean(loss)
    loss.backward()
    grads = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in m.named_parameters()}
    return el, pred.detach(), att, float(loss), grads


def main():
    asr_mod, charlm_mod = ref_shim.load(allow_container_reference=True)

    # ------------------------------------------------------------------ tiny (weights stored)
    dims = (50, 16, 16, 8, 12, 1.0)
    sd = O.make_state_dict(50, 16, 16, 8, 12, seed=1)
    m0 = ref_model(asr_mod, dims)
    for k, v in m0.state_dict().items():
        assert torch.equal(v, sd[k]), "make_state_dict diverges from reference init at " + k
    B, T, F, U = 5, 77, 12, 10
    g = torch.Generator().manual_seed(99)
    lens = [61, 53, 40, 33, 9]
    x = torch.randn(B, T, F, generator=g)
    for i, l in enumerate(lens):
        x[i, l:] = 0
    _, _, y = O.synth_batch(B, T, F, U, seed=5)
    el, pred, att, loss, grads = run_train(asr_mod, dims, sd, x, lens, y)
    # greedy (teacher=None) forward, the trainer.valid path (trainer.py:478)
    mg = ref_model(asr_mod, dims, sd)
    with torch.no_grad():
        _, gpred, gatt = mg(x, U + 3, state_len=lens)
    enc, _ = ref_model(asr_mod, dims, sd).encoder(x, lens)
    out = {'x': x.numpy(), 'lens': np.array(lens), 'y': y.numpy(), 'enc_len': np.array(el),
           'enc': enc.detach().numpy(), 'logits': pred.numpy(), 'att': att.numpy(), 'loss': np.float64(loss),
           'greedy_logits': gpred.numpy(), 'greedy_att': gatt.numpy(), 'dims': np.array(dims[:5])}
    for k, v in sd.items():
        out['sd.' + k] = v.numpy()
    for k, v in grads.items():
        out['grad.' + k] = v.numpy()
    # bs=1 decode with and without LM on the tiny model
    lm = O.make_charlm_state_dict(50, 128, seed=7)
    lmm = charlm_mod.CharLM(50, 128)
    lmm.load_state_dict(lm)
    lmm.eval()
    md = ref_model(asr_mod, dims, sd).eval()
    dec = []
    for i, l in enumerate(lens[:4]):
        xi = x[i:i + 1, :l]
        with torch.no_grad():
            s0 = md.decode(xi, [l], lmm, _Mapper(), 0.0)
            s1 = md.decode(xi, [l], lmm, _Mapper(), 0.5)
        dec.append((s0, s1))
    out['decode_lm0'] = np.array([d[0] for d in dec])
    out['decode_lm05'] = np.array([d[1] for d in dec])
    for k, v in lm.items():
        out['lm.' + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, 'las_tiny.npz'), **out)
    print('tiny: loss', loss, 'enc_len', el, 'decode', dec[0])

    # ------------------------------------------------------------------ default dims (weights from seed)
    dims = (50, 256, 256, 128, 80, 1.0)
    sd = O.make_state_dict(50, 256, 256, 128, 80, seed=1)
    m0 = ref_model(asr_mod, dims)
    for k, v in m0.state_dict().items():
        assert torch.equal(v, sd[k]), k
    B, T, F, U = 8, 128, 80, 20
    x, lens, y = O.synth_batch(B, T, F, U, seed=1234)
    el, pred, att, loss, grads = run_train(asr_mod, dims, sd, x, lens, y)
    enc, _ = ref_model(asr_mod, dims, sd).encoder(x, lens)
    out = {'B': B, 'T': T, 'F': F, 'U': U, 'enc_len': np.array(el), 'enc': enc.detach().numpy(),
           'logits': pred.numpy(), 'att': att.numpy(), 'loss': np.float64(loss)}
    for k, v in grads.items():
        out['gnorm.' + k] = np.float64(v.double().norm())
        out['ghead.' + k] = v.flatten()[:256].numpy()
    out['gnorm_total'] = np.float64(torch.sqrt(sum(v.double().pow(2).sum() for v in grads.values())))
    np.savez_compressed(os.path.join(HERE, 'las_default.npz'), **out)
    print('default: loss', loss, 'enc_len', el)

    # ------------------------------------------------------------------ decode, default dims, margin variant
    sdm = {k: v.clone() for k, v in sd.items()}
    sdm['char_trans.weight'] = sdm['char_trans.weight'] * 20.0      # SURVEY §8d C3 "margin" variant
    lmm = charlm_mod.CharLM(50, 128)
    lmm.load_state_dict(lm)
    lmm.eval()
    out = {}
    g = torch.Generator().manual_seed(4321)
    Ts = [int(v) for v in torch.randint(96, 161, (6,), generator=g)]
    for name, w in (('plain', sd), ('margin', sdm)):
        md = ref_model(asr_mod, dims, w).eval()
        for lmw in (0.0, 0.5):
            strs = []
            for i, Ti in enumerate(Ts):
                xi = torch.randn(1, Ti, 80, generator=torch.Generator().manual_seed(7000 + i))
                with torch.no_grad():
                    strs.append(md.decode(xi, [Ti], lmm, _Mapper(), lmw))
            out['%s_lm%s' % (name, str(lmw).replace('.', ''))] = np.array(strs)
            print(name, lmw, [s[:24] for s in strs[:2]], [len(s) for s in strs])
    out['Ts'] = np.array(Ts)
    np.savez_compressed(os.path.join(HERE, 'decode_default.npz'), **out)

    # ------------------------------------------------------------------ fbank (oracle output; reference unpinned)
    rng = np.random.RandomState(1234)
    y16 = (0.1 * rng.randn(16000)).astype(np.float32)
    out = {'y16': y16, 'fb16_80': FB.log_fbank(y16, 16000, 80), 'fb16_40': FB.log_fbank(y16, 16000, 40),
           'fb22_40': FB.log_fbank(y16[:11025], 22050, 40)}
    np.savez_compressed(os.path.join(HERE, 'fbank_1s.npz'), **out)
    print('fbank shapes', out['fb16_80'].shape, out['fb22_40'].shape)


if __name__ == '__main__':
    main()
