"""The oracle (oracle/*.py) against the committed outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py) -- CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import fbank_oracle as FB
from oracle import las_oracle as O
from oracle import las_port as P


def _rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope='module')
def tiny(golden_dir):
    z = np.load(os.path.join(golden_dir, 'las_tiny.npz'))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd.')}
    return z, sd


def test_state_dict_matches_reference_init(tiny):
    z, sd = tiny
    mine = O.make_state_dict(*[int(v) for v in z['dims']], seed=1)
    assert set(mine) == set(sd) and len(sd) == 46
    for k in sd:
        assert torch.equal(mine[k], sd[k]), k


def test_tiny_forward_loss_grads(tiny):
    z, sd = tiny
    x, lens, y = torch.from_numpy(z['x']), [int(v) for v in z['lens']], torch.from_numpy(z['y'])
    loss, logits, att, enc, grads = O.train_step_grads(sd, x, lens, y)
    assert np.abs(enc.numpy() - z['enc']).max() < 1e-5
    assert np.abs(logits.numpy() - z['logits']).max() < 1e-5          # fp32 path tolerance, SURVEY §8c
    assert np.abs(att.numpy() - z['att']).max() < 1e-6
    assert abs(float(loss) - float(z['loss'])) < 1e-6 * abs(float(z['loss'])) + 1e-6
    gtot = np.sqrt(sum(float((z['grad.' + k].astype(np.float64) ** 2).sum()) for k in sd))
    for k in sd:
        d = float((grads[k].double() - torch.from_numpy(z['grad.' + k]).double()).norm())
        assert d <= 1e-4 * float(np.linalg.norm(z['grad.' + k].astype(np.float64))) + 1e-6 * gtot, k


def test_tiny_fp64_oracle_close_to_fp32_reference(tiny):
    z, sd = tiny
    x, lens, y = torch.from_numpy(z['x']), [int(v) for v in z['lens']], torch.from_numpy(z['y'])
    loss, logits, *_ = O.train_step_grads(sd, x, lens, y, dtype=torch.float64)
    assert np.abs(logits.numpy() - z['logits']).max() < 1e-5
    assert abs(float(loss) - float(z['loss'])) < 1e-5


def test_tiny_greedy_forward(tiny):
    z, sd = tiny
    x, lens = torch.from_numpy(z['x']), [int(v) for v in z['lens']]
    U = z['greedy_logits'].shape[1]
    with torch.no_grad():
        el, logits, att, _ = O.asr_forward(sd, x, lens, U)
    assert el == [int(v) for v in z['enc_len']]
    assert np.array_equal(logits.argmax(-1).numpy(), z['greedy_logits'].argmax(-1))
    assert np.abs(logits.numpy() - z['greedy_logits']).max() < 1e-4


def test_tiny_decode_strings(tiny):
    z, sd = tiny
    lm = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('lm.')}
    x, lens = torch.from_numpy(z['x']), [int(v) for v in z['lens']]
    with torch.no_grad():
        for i in range(len(z['decode_lm0'])):
            xi = x[i:i + 1, :lens[i]]
            assert O.ids_to_str(O.decode_greedy(sd, xi, [lens[i]], lm, 0.0)) == str(z['decode_lm0'][i])
            ids, margin = O.decode_greedy(sd, xi, [lens[i]], lm, 0.5, return_margin=True)
            got, want = O.ids_to_str(ids), str(z['decode_lm05'][i])
            n = min(len(got), len(want), 20)
            assert got[:n] == want[:n] and (margin < 1e-4 or got == want)


def test_default_dims_forward_loss_grads(golden_dir):
    z = np.load(os.path.join(golden_dir, 'las_default.npz'))
    sd = O.make_state_dict(50, 256, 256, 128, 80, seed=1)
    x, lens, y = O.synth_batch(int(z['B']), int(z['T']), int(z['F']), int(z['U']), seed=1234)
    loss, logits, att, enc, grads = O.train_step_grads(sd, x, lens, y)
    assert np.abs(enc.numpy() - z['enc']).max() < 1e-5
    assert np.abs(logits.numpy() - z['logits']).max() < 1e-5
    assert np.abs(att.numpy() - z['att']).max() < 1e-6
    assert abs(float(loss) - float(z['loss'])) < 1e-6 * float(z['loss']) + 1e-6
    gtot = float(z['gnorm_total'])
    for k in sd:
        gn = float(grads[k].double().norm())
        assert abs(gn - float(z['gnorm.' + k])) <= 1e-4 * float(z['gnorm.' + k]) + 1e-6 * gtot, k
        head = grads[k].flatten()[:256].numpy()
        assert np.abs(head - z['ghead.' + k]).max() <= 1e-4 * np.abs(z['ghead.' + k]).max() + 1e-6 * gtot, k


def test_port_matches_oracle_and_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, 'las_default.npz'))
    sd = O.make_state_dict(50, 256, 256, 128, 80, seed=1)
    x, lens, y = O.synth_batch(int(z['B']), int(z['T']), int(z['F']), int(z['U']), seed=1234)
    port = P.Port(sd, tf_rate=1.0)
    loss, logits = port.train_step(x, lens, y)
    assert np.abs(logits.numpy() - z['logits']).max() < 1e-5
    assert abs(float(loss) - float(z['loss'])) < 1e-6 * float(z['loss']) + 1e-6
    g = port.grads()
    for k in sd:
        assert abs(float(g[k].double().norm()) - float(z['gnorm.' + k])) <= 1e-4 * float(z['gnorm.' + k]) + 1e-6 * float(z['gnorm_total']), k


def test_decode_default_margin_variant(golden_dir):
    z = np.load(os.path.join(golden_dir, 'decode_default.npz'))
    sd = O.make_state_dict(50, 256, 256, 128, 80, seed=1)
    sd['char_trans.weight'] = sd['char_trans.weight'] * 20.0
    lm = O.make_charlm_state_dict(50, 128, seed=7)
    port = P.Port(sd)
    with torch.no_grad():
        for i, Ti in enumerate([int(v) for v in z['Ts']][:3]):
            xi = torch.randn(1, Ti, 80, generator=torch.Generator().manual_seed(7000 + i))
            assert O.ids_to_str(O.decode_greedy(sd, xi, [Ti], None, 0.0)) == str(z['margin_lm00'][i])
            assert O.ids_to_str(O.decode_greedy(sd, xi, [Ti], lm, 0.5)) == str(z['margin_lm05'][i])
            assert O.ids_to_str(port.decode(xi, [Ti], lm, 0.5)) == str(z['margin_lm05'][i])


# ------------------------------------------------------------------------------------------ fbank
def test_fbank_golden_regression(golden_dir):
    z = np.load(os.path.join(golden_dir, 'fbank_1s.npz'))
    assert np.abs(FB.log_fbank(z['y16'], 16000, 80) - z['fb16_80']).max() < 1e-6
    assert np.abs(FB.log_fbank(z['y16'], 16000, 40) - z['fb16_40']).max() < 1e-6
    assert np.abs(FB.log_fbank(z['y16'][:11025], 22050, 40) - z['fb22_40']).max() < 1e-6
    assert z['fb16_80'].shape == (101, 80) and z['fb22_40'].shape == (51, 40)


def test_fbank_silence_and_shapes():
    out = FB.log_fbank(np.zeros(16000, np.float32), 16000, 80)
    assert out.dtype == np.float32 and out.shape == (101, 80)
    assert np.allclose(out, np.log(np.finfo(float).eps))
    for n in (400, 401, 1599, 1600, 1601, 160000):
        assert FB.log_fbank(np.ones(n, np.float32), 16000, 40).shape[0] == FB.num_frames(n, 16000) == 1 + n // 160
    assert FB.num_frames(22050, 22050) == 1 + (22050 - 1) // 220


def test_fbank_independent_witness_torchaudio(golden_dir):
    """librosa 0.6.3 is absent (parity unpinned by the reference); torchaudio's Slaney mel
    spectrogram is an independent implementation of the same published algorithm."""
    ta = pytest.importorskip('torchaudio')
    z = np.load(os.path.join(golden_dir, 'fbank_1s.npz'))
    y = torch.from_numpy(z['y16'])
    for n_mels in (40, 80):
        ms = ta.transforms.MelSpectrogram(sample_rate=16000, n_fft=400, win_length=400, hop_length=160, f_min=0.0,
                                          f_max=8000.0, n_mels=n_mels, window_fn=torch.hann_window, power=2.0,
                                          center=True, pad_mode='reflect', norm='slaney', mel_scale='slaney')
        w = torch.log(ms(y).double() + FB.EPS).t().numpy()
        assert np.abs(w - z['fb16_%d' % n_mels]).max() < 5e-4


def test_fbank_independent_witness_transformers(golden_dir):
    au = pytest.importorskip('transformers.audio_utils')
    z = np.load(os.path.join(golden_dir, 'fbank_1s.npz'))
    mel = au.mel_filter_bank(num_frequency_bins=201, num_mel_filters=80, min_frequency=0.0, max_frequency=8000.0,
                             sampling_rate=16000, norm='slaney', mel_scale='slaney')
    spec = au.spectrogram(z['y16'].astype(np.float64), au.window_function(400, 'hann', periodic=True), frame_length=400,
                          hop_length=160, fft_length=400, power=2.0, center=True, pad_mode='reflect', mel_filters=mel)
    w = np.log(spec + FB.EPS).T
    assert np.abs(w - z['fb16_80']).max() < 1e-4


# ------------------------------------------------------------------------------------------ live
def test_oracle_vs_reference_live():
    """Only where /root/reference exists: a second, differently shaped case straight against the
    unmodified reference (odd T at every layer, length-1 encoder rows)."""
    from oracle import ref_shim
    if not ref_shim.available(allow_container_reference=True):
        pytest.skip('reference sources not present on this box')
    asr_mod, _ = ref_shim.load(allow_container_reference=True)
    dims = (50, 24, 16, 8, 10, 1.0)
    sd = O.make_state_dict(50, 24, 16, 8, 10, seed=3)
    m = asr_mod.ASR(*dims)
    m.load_state_dict(sd)
    B, T, U = 6, 91, 7
    g = torch.Generator().manual_seed(11)
    lens = [91, 90, 75, 50, 23, 8]
    x = torch.randn(B, T, 10, generator=g)
    for i, l in enumerate(lens):
        x[i, l:] = 0
    y = torch.randint(3, 50, (B, U + 2), generator=g)
    y[:, 0] = 0
    y[:, U + 1] = 1
    el, pred, att = m(x, U + 1, teacher=y, state_len=lens)
    with torch.no_grad():
        el2, logits, att2, _ = O.asr_forward(sd, x, lens, U + 1, teacher=y)
    assert el == el2
    assert float((pred - logits).abs().max()) < 1e-5 and float((att - att2).abs().max()) < 1e-6


# ------------------------------------------------------------------------------------------------ validation metrics
def test_postprocess_oracle_matches_reference_outputs(golden_dir):
    """oracle/postprocess_oracle.py against postprocess.calc_acc / calc_err + ASRDataset.Mapper of the unmodified reference
    (tests/golden/postprocess.npz): Python-float results identical, hypotheses string-identical."""
    from oracle import postprocess_oracle as PO
    z = np.load(os.path.join(golden_dir, 'postprocess.npz'))
    k = 0
    while 'case_%d' % k in z.files:
        seed, B, U, L = (int(v) for v in z['case_%d' % k])
        predict, label, pred_tok = PO.synth_cases(seed=seed, B=B, U=U, L=L)
        assert np.array_equal(PO.argmax_tokens(predict), pred_tok)                 # ties resolve to the first maximum
        assert PO.calc_acc(predict, label) == float(z['acc_%d' % k])
        assert PO.calc_err(predict, label) == float(z['err_%d' % k])
        st = PO.utterance_stats(predict, label)
        assert np.array_equal(st[:, 0] / st[:, 1], z['acc_utt_%d' % k])
        assert np.array_equal(st[:, 2] / st[:, 3], z['err_utt_%d' % k])
        assert [PO.translate(p) for p in pred_tok] == [str(s) for s in z['hyp_%d' % k]]
        k += 1
    assert k == 4


def test_postprocess_oracle_known_answers():
    from oracle import postprocess_oracle as PO
    assert PO.levenshtein('kitten', 'sitting') == 3 and PO.levenshtein([], ['a']) == 1 and PO.levenshtein(['a'], ['a']) == 0
    assert PO.translate([3, 46, 46, 4, 1, 5]) == 'a  b' and PO.translate([0, 1]) == '' and PO.translate([3, 0, 4]) == 'ab'
    assert ''.split(' ') == [''] and 'a  b'.split(' ') == ['a', '', 'b']


def test_postprocess_oracle_vs_reference_live(golden_dir):
    from oracle import ref_shim
    from oracle import postprocess_oracle as PO
    if not ref_shim.available(allow_container_reference=True):
        pytest.skip('reference sources not present on this box')
    ref_shim.load(allow_container_reference=True)
    import ASRDataset
    import postprocess
    mapper = ASRDataset.Mapper()
    for seed in (11, 12):
        predict, label, _ = PO.synth_cases(seed=seed, B=12, U=23, L=14)
        pt, lt = torch.from_numpy(predict), torch.from_numpy(label)
        assert postprocess.calc_acc(pt, lt) == PO.calc_acc(predict, label)
        assert postprocess.calc_err(pt, lt, mapper) == PO.calc_err(predict, label)


def test_oracle_beam_search_reduces_to_greedy():
    """`O.decode_beam` (the checker of ASR.beam_decode_batch; the reference has no beam search, trainer.py:590) with beam size 1 is
    `O.decode_greedy`, itself pinned to the reference's strings above: same tokens with and without the CharLM; wider beams return
    a hypothesis whose score is the sum of its own per-step scores (re-scored independently)."""
    import torch
    dims = (50, 32, 32, 16, 20)
    sd = O.make_state_dict(*dims, seed=4)
    sd['char_trans.weight'] = sd['char_trans.weight'] * 8.0
    lm = O.make_charlm_state_dict(50, 128, seed=7)
    x, lens, _ = O.synth_batch(4, 48, 20, 6, seed=9)
    for i in range(4):
        t = lens[i]
        xi = x[i:i + 1, :t]
        for kw in ({}, {'lm': lm, 'lm_weight': 0.5}):
            assert O.decode_beam(sd, xi, [t], 1, max_steps=12, **kw) == O.decode_greedy(sd, xi, [t], max_steps=12, **kw)
        ids, score = O.decode_beam(sd, xi, [t], 4, max_steps=12, return_score=True)
        assert len(ids) <= 12 and score <= 0.0
