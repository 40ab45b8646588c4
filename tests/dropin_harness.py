"""Test harness: drives the UNMODIFIED reference trainers (ASRTrainer.exec / ASRTester.exec, trainer.py:372-592, through
oracle/ref_shim.py) on a synthetic `.npy` + `index.tsv` dataset in the reference's own on-disk format
(preprocess.py:47-60,253-269), either with the reference's `asr` module or with `ss_asr_b200.asr` swapped in under the
module name `asr` -- the two-line integration INTEGRATION.md describes.  Test infrastructure only."""
import argparse
import copy
import importlib
import os
import random
import sys

import numpy as np
import torch
import yaml

from oracle import ref_shim

CHARS = 'abdefghijklmnoprstuvxyáéíóúýæöþð '      # a subset of preprocess.py:17-19 (no c, q, w, z in the reference's alphabet)


def make_dataset(root, n_utt=8, feat=40, t_min=48, t_max=96, seed=0, descending_batches=4):
    """Writes fbanks/<id>.npy (float64, zero-padded to the longest utterance, preprocess.py:253-269) and index.tsv with the six
    columns of preprocess.py:52-56.  Rows are ordered so that every batch of `descending_batches` utterances has decreasing
    frame counts (pack_padded_sequence contract, conf/README.md:16)."""
    rng = np.random.RandomState(seed)
    os.makedirs(os.path.join(root, 'fbanks'), exist_ok=True)
    lens = sorted((int(v) for v in rng.randint(t_min, t_max + 1, n_utt)), reverse=True)
    max_len = max(lens)
    rows = []
    for i, n in enumerate(lens):
        fb = np.zeros((max_len, feat))
        fb[:n] = rng.randn(n, feat)
        path = os.path.join(root, 'fbanks', 'utt%03d.npy' % i)
        np.save(path, fb)
        text = ''.join(CHARS[int(c)] for c in rng.randint(0, len(CHARS), rng.randint(5, 12))).strip() or 'a'
        text = ' '.join(text.split())
        rows.append(('<' + text + '>', path, len(text) + 2, n, 'utt%03d.txt' % i, 'utt%03d.wav' % i))
    # batches of decreasing length: rows are globally descending already
    index = os.path.join(root, 'index.tsv')
    with open(index, 'w', encoding='utf-8') as f:
        for r in rows:
            f.write('\t'.join(str(a) for a in r) + '\n')
    return index


def load_config(index, batch_size=4, allow_container_reference=False):
    src = ref_shim.ref_src(allow_container_reference)
    conf = os.path.join(os.path.dirname(src), 'conf', 'default.yaml')
    cfg = yaml.safe_load(open(conf))
    a = cfg['asr']
    a['train_index'] = a['valid_index'] = a['test_index'] = index
    a['mdl'] = dict(a['mdl'], tf_rate=1.0)        # deterministic teacher forcing: sampled steps use different RNGs by design
    a.update(n_epochs=1, train_batch_size=batch_size, valid_batch_size=batch_size, logging_step=1, wer_step=1,
             valid_step=10 ** 6, save_step=10 ** 6)
    cfg['char_lm']['hidden_size'] = cfg['char_lm']['mdl']['hidden_size']        # trainer.py:568 vs default.yaml:88-89
    return cfg


def import_trainer(ours, allow_container_reference=False):
    """The reference's `trainer` module, imported with `asr` resolving to the reference (ours=False) or to ss_asr_b200.asr
    (ours=True).  Returns (module, restore) -- call restore() to put sys.modules back."""
    ref_shim.load(allow_container_reference)
    saved = {k: sys.modules.get(k) for k in ('asr', 'trainer')}
    sys.modules.pop('trainer', None)
    if ours:
        import ss_asr_b200.asr as mine
        sys.modules['asr'] = mine
    with ref_shim.stubs():
        tr = importlib.import_module('trainer')

    def restore():
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return tr, restore


def _paras(workdir, name):
    return argparse.Namespace(ckpdir=os.path.join(workdir, 'ckpt'), logdir=os.path.join(workdir, 'log'), name=name,
                              verbose=False, gpu=False, seed=1)


def _seed(seed=1):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def run_trainer(tr, cfg, workdir, name, device):
    """train.py:54-71 for ASRTrainer: seeds, load_data, set_model, exec (one epoch = 2 steps of 4 utterances).  The validation
    pass that the reference runs at step 0 is executed too (trainer.py:453-455); its NameError defect (trainer.py:530-532,
    after the metrics have been logged) is caught.  Returns (state_dict on the CPU, logged scalars)."""
    _seed(1)
    s = tr.ASRTrainer(copy.deepcopy(cfg), _paras(workdir, name))
    s.device = torch.device(device)
    s.paras.gpu = s.device.type == 'cuda'
    s.load_data()
    s.set_model()
    orig_valid = s.valid

    def valid():
        try:
            orig_valid()
        except NameError:
            pass
        s.asr_model.train()
    s.valid = valid
    s.exec()
    sd = {k: v.detach().cpu().clone() for k, v in s.asr_model.state_dict().items()}
    return sd, list(s.lg.log.rec)


def run_tester(tr, cfg, workdir, name, device, state_dict):
    """train.py:54-71 for ASRTester with a checkpoint at the path ASRTester loads from (trainer.py:563-565)."""
    p = _paras(workdir, name)
    os.makedirs(os.path.join(p.ckpdir, name), exist_ok=True)
    torch.save(state_dict, os.path.join(p.ckpdir, name, 'asr.cpt'))
    _seed(7)
    s = tr.ASRTester(copy.deepcopy(cfg), p)
    s.device = torch.device(device)
    s.paras.gpu = s.device.type == 'cuda'
    s.load_data()
    s.set_model()
    return s.exec()
