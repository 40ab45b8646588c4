"""CPU-only: the C-ABI library builds, loads and exports every symbol include/ssasr.h declares; host-side
module surface mirrors the reference (state_dict keys/shapes, init stream)."""
import os
import re

import torch

from oracle import las_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ss_asr_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, 'include', 'ssasr.h')).read()
    declared = set(re.findall(r'\b(ssasr_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.ssasr_fbank_num_frames(160000, 16000) == 1001
    assert lib.ssasr_fbank_num_frames(22050, 22050) == 1 + (22050 - 1) // 220


def test_module_surface_matches_reference():
    from ss_asr_b200.asr import ASR, Listener
    torch.manual_seed(1)
    m = ASR(50, 256, 256, 128, 40, 0.9)
    sd = m.state_dict()
    assert len(sd) == 46 and sum(v.numel() for v in sd.values()) == 10187954
    ref = O.make_state_dict(50, 256, 256, 128, 40, seed=1)
    assert set(sd) == set(ref)
    for k in sd:
        assert torch.equal(sd[k], ref[k]), k
    assert m.decoder.layer_1.bias_ih[256:512].eq(1).all() and m.tf_rate == 0.9
    l = Listener(256, 40)
    assert l.get_outdim() == 512 and l.out_dim == 512
    m.decoder.init_rnn(3, torch.device('cpu'))
    assert len(m.decoder.state_list) == 2 and m.decoder.state_list[0].shape == (3, 256)
    h, c = m.decoder.hidden_state
    assert len(h) == 2 and len(c) == 2
    m.attention.reset_enc_mem()
    assert m.attention.comp_listener_feature is None


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'ss_asr_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+\.*oracle', src, re.M), f
