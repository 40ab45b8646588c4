"""Drop-in Listen-Attend-Spell modules for the ss_asr trainers, backed by libssasr.so (sm_100a CUDA).

Mirrors the public surface of /root/reference/src/asr.py -- same class names, constructor signatures,
attribute names, `state_dict` keys/shapes and return tuples -- so `trainer.py`'s
ASRTrainer / ASRTester call sites (trainer.py:397-398, 422, 478, 591) work unchanged after
`from ss_asr_b200.asr import ASR`.  The torch.nn containers below (nn.LSTM, nn.LSTMCell, nn.Linear,
nn.Embedding) are used ONLY as parameter holders, which gives identical initialisation order, strict
`load_state_dict` compatibility and `.to(device)` behaviour; their own forward() is never called.
"""
import math
import random

import numpy as np

import torch
import torch.nn as nn

from . import functional as Fk

EOS_TKN = '>'          # preprocess.py:25


def _lens_list(state_len):
    if torch.is_tensor(state_len):
        state_len = state_len.tolist()
    return [int(s) for s in state_len]


def _i32_dev(vals, device):
    """Python ints -> int32 device tensor WITHOUT blocking the host.  `torch.tensor(list, device='cuda')` copies from pageable
    memory: the host then waits until the stream has drained, once per layer, and the GPU idles while the next layer is being
    enqueued (measured: 1.25 ms per call, 5 ms of an 10 ms C4 step).  Staged in pinned memory (torch's caching host allocator
    keeps the block alive until the copy has run) and copied asynchronously on the current stream instead."""
    h = torch.empty(len(vals), dtype=torch.int32, pin_memory=True)
    h.numpy()[:] = vals
    return h.to(device, non_blocking=True)


def _blstm_params(lstm):
    return (lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0,
            lstm.weight_ih_l0_reverse, lstm.weight_hh_l0_reverse, lstm.bias_ih_l0_reverse, lstm.bias_hh_l0_reverse)


def _fit_time(x, t_h):
    """Slices / zero-pads the time axis of x [B,T,K] to exactly t_h frames."""
    T = x.shape[1]
    if T == t_h:
        return x
    if T > t_h:
        return x[:, :t_h]
    return torch.nn.functional.pad(x, (0, 0, 0, t_h - T))


class pBLSTM(nn.Module):
    """asr.py:394-450: BLSTM over a (packed) batch followed by frame-pair concatenation."""

    def __init__(self, in_dim, out_dim):
        super(pBLSTM, self).__init__()
        self.layer = nn.LSTM(in_dim, out_dim, bidirectional=True, batch_first=True)   # parameter holder
        self.precision = 'fp32'
        self.bf16_input = None      # set by Listener: bf16 twin of the next forward's input (consumed by that call)
        self.bf16_output = None     # bf16 twin of the last forward's output, when the recurrent kernel produced one

    def forward(self, input_x, state=None, state_len=None, pack_input=False):
        if state is not None:
            raise NotImplementedError('pBLSTM: a non-zero initial state is never passed on the ASR path '
                                      '(asr.py:256-261) and is not supported by the CUDA kernels')
        B, T, _ = input_x.shape
        if pack_input:
            assert state_len is not None, "Please specify seq len for pack_padded_sequence."
            lens = _lens_list(state_len)
            if any(lens[i] < lens[i + 1] for i in range(len(lens) - 1)) or lens[-1] <= 0:
                raise RuntimeError('pBLSTM: `state_len` must be sorted in decreasing order and positive '
                                   '(pack_padded_sequence contract, asr.py:413)')
            run_lens = lens
            t_max = lens[0]
        else:
            run_lens = [T] * B
            t_max = T
        if t_max > T:
            raise RuntimeError('pBLSTM: length %d exceeds the padded time dimension %d' % (t_max, T))
        t_h = t_max + (t_max & 1)
        x = _fit_time(input_x, t_h)
        lens_dev = _i32_dev(run_lens, input_x.device)
        if self.bf16_input is not None and x is input_x:
            Fk.hint_bf16_input(self.bf16_input)       # the previous layer's bf16 output (tensor-core path): no conversion pass
        self.bf16_input = None
        hout = Fk.blstm(x, lens_dev, True, _blstm_params(self.layer), self.precision)   # [B, t_h, 2S]
        out = hout.view(B, t_h // 2, 2 * hout.shape[2])[:, :t_max // 2]              # downsample = a view
        hb = Fk.LAST_BLSTM['hb']
        # bf16 twin of `out`, valid only when the view above drops nothing
        self.bf16_output = hb.view(B, t_h // 2, 2 * hout.shape[2]) if (hb is not None and t_max == t_h) else None
        Fk.LAST_BLSTM['hb'] = None
        hidden = None   # (h_n, c_n) is discarded by every caller on the ASR path (asr.py:256-262)
        if state_len is not None:
            if pack_input:
                new_len = [int(s / 2) for s in run_lens]
            else:
                new_len = [int(s / 2) for s in _lens_list(state_len)]
            return out, hidden, new_len
        return out, hidden

    def downsample(self, x):
        t_dim, f_dim = x.shape[1], x.shape[2]
        t2 = t_dim // 2
        return x[:, :2 * t2, :].contiguous().view(-1, t2, f_dim * 2)


class Listener(nn.Module):
    """asr.py:214-264."""

    def __init__(self, state_size, feature_dim):
        super(Listener, self).__init__()
        self.state_size = state_size
        self.out_dim = 2 * self.state_size
        self.blstm_1 = pBLSTM(feature_dim, self.state_size)
        self.blstm_2 = pBLSTM(self.state_size * 2 * 2, self.state_size)
        self.blstm_3 = pBLSTM(self.state_size * 2 * 2, self.state_size)
        self.blstm_4 = nn.LSTM(self.state_size * 2 * 2, self.state_size, bidirectional=True)   # parameter holder
        self.utterance_independent = False   # True: run blstm_4 as bs=1 would (ASR.decode batching)
        self.precision = 'fp32'

    def set_precision(self, precision):
        """'fp32': exact SIMT path (decode / validation / tight parity); 'bf16': tcgen05 gate GEMMs (training)."""
        assert precision in ('fp32', 'bf16', 'tf32x3')
        self.precision = precision
        for m in (self.blstm_1, self.blstm_2, self.blstm_3):
            m.precision = precision

    def get_outdim(self):
        return self.out_dim

    def forward(self, x, state_len, pack_input=True):
        t_in = x.shape[1]
        x, _, state_len = self.blstm_1(x, state_len=state_len, pack_input=pack_input)
        self.blstm_2.bf16_input = self.blstm_1.bf16_output
        x, _, state_len = self.blstm_2(x, state_len=state_len, pack_input=pack_input)
        self.blstm_3.bf16_input = self.blstm_2.bf16_output
        x, _, state_len = self.blstm_3(x, state_len=state_len, pack_input=pack_input)
        x_bf = self.blstm_3.bf16_output
        for m in (self.blstm_1, self.blstm_2, self.blstm_3):
            m.bf16_output = None
        if x.shape[1] == 0:
            raise RuntimeError('Listener: at least 8 input frames are needed (three frame-pair reductions leave none of %d); '
                               'the reference fails in nn.LSTM at the same point (asr.py:262)' % int(t_in))
        x = x.contiguous()
        if self.utterance_independent:
            # bs=1 semantics for every utterance at once: one cell step from zero state per frame
            B, Tp, K = x.shape
            x = Fk.blstm(x.view(1, B * Tp, K), None, False, _blstm_params(self.blstm_4), self.precision).view(B, Tp, -1)
        else:
            # seq-first quirk (asr.py:237-238,262): dim 0 (utterances) is the time axis of blstm_4
            if x_bf is not None and x_bf.shape == x.shape:
                Fk.hint_bf16_input(x_bf)
            x = Fk.blstm(x, None, False, _blstm_params(self.blstm_4), self.precision)
        return x, state_len


class Speller(nn.Module):
    """asr.py:267-326 (parameter holder + per-utterance state attributes)."""

    def __init__(self, state_size, encoder_out_size):
        super(Speller, self).__init__()
        self.layer_1 = nn.LSTMCell(input_size=encoder_out_size + state_size, hidden_size=state_size)
        self.layer_2 = nn.LSTMCell(input_size=state_size, hidden_size=state_size)
        self.state_list = []
        self.cell_list = []
        self.state_size = state_size
        self.num_layers = 2

    def init_rnn(self, batch_size, device):
        self.state_list = [torch.zeros(batch_size, self.state_size).to(device)] * self.num_layers
        self.cell_list = [torch.zeros(batch_size, self.state_size).to(device)] * self.num_layers

    @property
    def hidden_state(self):
        return [s.clone().detach().cpu() for s in self.state_list], \
            [c.clone().detach().cpu() for c in self.cell_list]

    @hidden_state.setter
    def hidden_state(self, state):
        device = self.state_list[0].device
        self.state_list = [s.to(device) for s in state[0]]
        self.cell_list = [c.to(device) for c in state[1]]

    def forward(self, input_context):
        """One decoding step of both cells (asr.py:314-326); state lives on self.state_list / self.cell_list."""
        self.state_list[0], self.cell_list[0] = Fk.lstm_cell(input_context, self.state_list[0], self.cell_list[0], self.layer_1)
        self.state_list[1], self.cell_list[1] = Fk.lstm_cell(self.state_list[0], self.state_list[1], self.cell_list[1], self.layer_2)
        return self.state_list[-1]

    def params(self):
        l1, l2 = self.layer_1, self.layer_2
        return (l1.weight_ih, l1.weight_hh, l1.bias_ih, l1.bias_hh, l2.weight_ih, l2.weight_hh, l2.bias_ih, l2.bias_hh)


class Attention(nn.Module):
    """asr.py:328-392 (parameter holder + cached-memory attributes)."""

    def __init__(self, mlp_out_size, encoder_out_size, decoder_state_size):
        super(Attention, self).__init__()
        self.softmax = nn.Softmax(dim=-1)
        self.phi = nn.Linear(decoder_state_size, mlp_out_size, bias=False)
        self.psi = nn.Linear(encoder_out_size, mlp_out_size)
        self.comp_listener_feature = None
        self.state_mask = None
        self._lens_dev = None

    def reset_enc_mem(self):
        self.comp_listener_feature = None
        self.state_mask = None
        self._lens_dev = None

    def forward(self, decoder_state, listener_feature, state_len):
        """-> (attention_score [B,T'], context [B,E])   asr.py:343-392 (memory and pad mask cached across steps)."""
        if self.comp_listener_feature is None:
            B, Tp = listener_feature.shape[0], listener_feature.shape[1]
            lens = _lens_list(state_len)
            self._lens_dev = _i32_dev(lens, listener_feature.device)
            self.state_mask = torch.arange(Tp, device=listener_feature.device)[None, :] >= self._lens_dev[:, None]
            self.comp_listener_feature = Fk.psi_memory(listener_feature, self.psi.weight, self.psi.bias)
        return Fk.attn_step(decoder_state, listener_feature, self.comp_listener_feature, self._lens_dev, self.phi.weight)

    def params(self):
        return (self.phi.weight, self.psi.weight, self.psi.bias)


class ASR(nn.Module):
    """asr.py:15-212."""

    def __init__(self, output_dim, encoder_state_size, decoder_state_size, mlp_out_size, feature_dim, tf_rate):
        super(ASR, self).__init__()
        enc_out_dim = encoder_state_size * 2
        self.encoder = Listener(encoder_state_size, feature_dim)
        self.attention = Attention(mlp_out_size, enc_out_dim, decoder_state_size)
        self.decoder = Speller(decoder_state_size, enc_out_dim)
        self.embed = nn.Embedding(output_dim, decoder_state_size)
        self.char_trans = nn.Linear(decoder_state_size, output_dim)
        self.tf_rate = tf_rate
        self.att_on_device = False       # True: keep attention maps on the GPU (skips the reference's D2H copy)
        self.att_async = False           # True: the D2H copy of the attention maps goes to a pinned buffer without blocking the
                                         # host (valid after the next stream sync, e.g. the loss read of trainer.py:440);
                                         # False: the reference's blocking .cpu() (asr.py:104)
        self._att_pinned = None
        self.train_precision = 'fp32'    # 'bf16': tensor-core gate GEMMs when training with grad enabled
        self.decode_precision = 'fp32'   # 'tf32x3': encoder input projections of decode_batch on tensor cores
        self.sample_seed = 0
        self.last_tokens = None          # [B,U] int32: the input token of every step of the last forward
        self.decode_stop_check = 16      # decode_batch: steps between 'has every utterance emitted EOS?' checks (0 = never)
        self.last_decode_steps = 0
        self.decode_encoder_chunk = 512  # decode_batch: utterances per Listener pass (0 = one pass over the whole batch)
        self.init_parameters()

    # ------------------------------------------------------------------------------------------
    def _spell(self, enc, enc_len, tok_in, modes, precision='fp32', lm=None, need_logits=True, stop_every=0, stop_token=1):
        lens_dev = _i32_dev(enc_len, enc.device)
        params = self.attention.params() + self.decoder.params() + (self.embed.weight, self.char_trans.weight,
                                                                    self.char_trans.bias)
        self.sample_seed += 1
        return Fk.spell(enc, lens_dev, tok_in, modes, self.sample_seed, params, precision, lm, need_logits,
                        stop_token=int(stop_token), stop_every=stop_every)

    def forward(self, audio_feature, decode_step, teacher=None, state_len=None):
        """-> (encode_len, logits [B,U,C] on the device, attention maps [B,U,T'] on the CPU)   asr.py:52-110"""
        Fk.join_deferred()
        use_bf16 = self.train_precision == 'bf16' and self.training and torch.is_grad_enabled() and teacher is not None
        self.encoder.set_precision('bf16' if use_bf16 else 'fp32')
        try:
            encode_feature, encode_len = self.encoder(audio_feature, state_len)
        finally:
            self.encoder.set_precision('fp32')
        B = audio_feature.shape[0]
        U = int(decode_step)
        tok_in = torch.zeros(B, U, dtype=torch.int32, device=encode_feature.device)
        if teacher is not None:
            n = min(U, teacher.shape[1])
            tok_in[:, 1:n] = teacher[:, 1:n].to(torch.int32)
            # one host draw per step, same stream of draws as asr.py:94
            modes = [0 if random.random() <= self.tf_rate else 2 for _ in range(U)]
        else:
            modes = [1] * U
        logits, att, toks = self._spell(encode_feature, encode_len, tok_in, modes, 'bf16' if use_bf16 else 'fp32')
        self.last_tokens = toks
        att = att.detach()
        if self.att_on_device:
            return encode_len, logits, att
        if self.att_async:
            if self._att_pinned is None or self._att_pinned.shape != att.shape:
                self._att_pinned = torch.empty(att.shape, dtype=att.dtype, pin_memory=True)
            self._att_pinned.copy_(att, non_blocking=True)
            return encode_len, logits, self._att_pinned
        return encode_len, logits, att.cpu()

    @torch.no_grad()
    def _encode_for_decode(self, xs, x_lens, prec):
        """Listener pass of the decoders with the per-utterance (bs=1) semantics of ASR.decode -> (enc [N,T',E], enc_len)."""
        prev = self.encoder.utterance_independent
        self.encoder.utterance_independent = True
        self.encoder.set_precision(prec)
        N = xs.shape[0]
        lens = _lens_list(x_lens)
        chunk = int(self.decode_encoder_chunk or 0)
        fetch = None
        if not xs.is_cuda:
            # HOST batch (pinned for an asynchronous copy): every Listener group is uploaded on a copy stream, only as many frames
            # deep as its longest utterance, the next group's copy running under the current group's encoder pass
            dev = self.embed.weight.device
            if dev.type != 'cuda':
                raise RuntimeError('ss_asr_b200: the model must live on a CUDA device (there is no CPU fallback)')
            main = torch.cuda.current_stream(dev)
            if getattr(self, '_copy_stream', None) is None:
                self._copy_stream = torch.cuda.Stream(dev)
            cs = self._copy_stream

            def fetch(r0, r1, tmax):
                cs.wait_stream(main)
                with torch.cuda.stream(cs):
                    src = xs[r0:r1, :tmax]
                    if src.is_contiguous() or xs.dtype != torch.float32 or not xs.is_contiguous():
                        t = src.to(dev, non_blocking=True)
                    else:
                        # rows of tmax frames out of the T-frame padded batch: one strided copy (torch would stage a
                        # contiguous host temporary first)
                        from . import _lib
                        t = torch.empty(r1 - r0, tmax, xs.shape[2], device=dev)
                        _lib.check(_lib.load().ssasr_memcpy2d_h2d(t.data_ptr(), tmax * xs.shape[2] * 4, src.data_ptr(),
                                                                  xs.shape[1] * xs.shape[2] * 4, tmax * xs.shape[2] * 4, r1 - r0,
                                                                  cs.cuda_stream), 'ssasr_memcpy2d_h2d')
                    ev = torch.cuda.Event()
                    ev.record(cs)
                t.record_stream(main)
                return t, ev
        try:
            if chunk <= 0 or N <= chunk:
                if fetch is not None:
                    xs, ev = fetch(0, N, xs.shape[1])
                    main.wait_event(ev)
                enc, enc_len = self.encoder(xs, x_lens)
            else:
                # utterances are independent and sorted by length: the Listener runs over groups of `chunk` utterances, each
                # only as many frames deep as ITS longest utterance (the recurrent kernels are bound by dependent steps, so
                # the padded tail of a short group is pure waste)
                enc, enc_len = None, []
                bounds = [(r0, min(N, r0 + chunk)) for r0 in range(0, N, chunk)]
                nxt = fetch(bounds[0][0], bounds[0][1], lens[bounds[0][0]]) if fetch is not None else None
                for i, (r0, r1) in enumerate(bounds):
                    if fetch is not None:
                        xg, ev = nxt
                        if i + 1 < len(bounds):
                            nxt = fetch(bounds[i + 1][0], bounds[i + 1][1], lens[bounds[i + 1][0]])
                        main.wait_event(ev)
                    else:
                        xg = xs[r0:r1, :lens[r0]]
                    e, el = self.encoder(xg, lens[r0:r1])
                    if enc is None:
                        enc = torch.zeros(N, e.shape[1], e.shape[2], dtype=e.dtype, device=e.device)
                    enc[r0:r1, :e.shape[1]] = e
                    enc_len += list(el)
        finally:
            self.encoder.utterance_independent = prev
            self.encoder.set_precision('fp32')
        return enc, enc_len

    @torch.no_grad()
    def decode_batch(self, xs, x_lens, max_steps=200, rnn_lm=None, lm_weight=0.0, precision=None, eos_id=1):
        """Greedy decoding of many utterances at once with the per-utterance (bs=1) semantics of ASR.decode:
        xs [N,T,F] zero-padded (on the device, or a host tensor -- pinned for asynchronous, pipelined uploads), x_lens sorted in
        decreasing order; optional CharLM rescoring (asr.py:153-159).
        `eos_id`: the token that ends an utterance (`mapper.char_to_ind(EOS_TKN)`, asr.py:167; 1 with the default Mapper).
        Returns a list of token-id lists."""
        # 'fp32': SIMT input projections (bit-for-tolerance exact path); 'tf32x3': the same math on tensor cores with split
        # operands (x = hi + lo in bf16, three products; 4e-7 .. 1.3e-6 from the fp32 path on the encoder states)
        prec = precision or self.decode_precision
        N = xs.shape[0]
        enc, enc_len = self._encode_for_decode(xs, x_lens, prec)
        tok_in = torch.zeros(N, max_steps + 1, dtype=torch.int32, device=enc.device)
        lm = None
        if rnn_lm is not None and lm_weight != 0:
            lm = (Fk.pack_charlm(rnn_lm, enc.device), lm_weight)
        # the loop stops as soon as every utterance has emitted EOS (checked every `decode_stop_check` steps; asr.py:161-162)
        _, _, toks = self._spell(enc, enc_len, tok_in, [3 if lm is not None else 1] * (max_steps + 1), precision=prec, lm=lm,
                                 need_logits=False, stop_every=int(self.decode_stop_check or 0), stop_token=eos_id)
        self.last_decode_steps = Fk.LAST_SPELL['steps_run']
        # every utterance's tokens up to (not including) its first EOS: vectorised on the host (the per-token Python loop took
        # 7 ms per 1000 utterances, a fifth of the whole call)
        t = toks[:, 1:max_steps + 1].cpu().numpy()
        is_eos = t == eos_id
        n = np.where(is_eos.any(1), is_eos.argmax(1), t.shape[1])
        return [t[i, :n[i]].tolist() for i in range(t.shape[0])]

    def decode(self, x, x_len, rnn_lm, mapper, lm_weight):
        """asr.py:112-173 (bs=1).  lm_weight == 0 runs entirely in the fused kernels."""
        assert len(x.shape) == 3 and x.shape[0] == 1
        ids = self.decode_batch(x, x_len, rnn_lm=rnn_lm, lm_weight=lm_weight, eos_id=int(mapper.char_to_ind(EOS_TKN)))[0]
        return ''.join(mapper.ind_to_char(i) for i in ids)

    @torch.no_grad()
    def beam_decode_batch(self, xs, x_lens, beam_size, max_steps=200, rnn_lm=None, lm_weight=0.0, precision=None, eos_id=1,
                          return_scores=False):
        """Beam search over the step of ASR.decode (asr.py:143-172) for many utterances at once, per-utterance (bs=1) semantics.
        The reference configures a beam (`decode_beam_size`, conf/default.yaml:16, trainer.py:552-554) but decodes greedily
        (trainer.py:590 TODO); the semantics are those of `oracle/las_oracle.py:decode_beam` (the checker) and reduce to
        `decode_batch` at beam_size 1: hypothesis score = sum of `log_softmax(asr) + lm_weight * log_softmax(lm)` over its tokens,
        finished hypotheses stay candidates, ties -> lower (parent, token), the best survivor is returned.
        The W hypotheses of utterance n are rows n W .. n W + W - 1 of every state tensor; per step: attention + both cells on the
        per-step kernels (query = layer-1 state of the previous step), character projection, `ssasr_beam_select` (scores, top-W,
        parents), `ssasr_gather_rows` (states by parent, embeddings by token); (parent, token) per step are back-tracked on the
        host at the end.  `rnn_lm`: the reference's CharLM module, called batched over the hypotheses as asr.py:154 calls it.
        Returns a list of token-id lists (and the scores with return_scores)."""
        from . import _lib
        lib = _lib.load()
        W = int(beam_size)
        assert 1 <= W <= 16, 'beam_size 1..16'
        prec = precision or self.decode_precision
        N = xs.shape[0]
        enc, enc_len = self._encode_for_decode(xs, x_lens, prec)
        dev = enc.device
        NW = N * W
        C, Sd = self.char_trans.weight.shape
        with torch.cuda.device(dev):
            st = _lib.stream()
            psi = Fk.psi_memory(enc, self.attention.psi.weight, self.attention.psi.bias)
            enc_r = enc.repeat_interleave(W, 0).contiguous()                  # hypotheses of an utterance share its memory
            psi_r = psi.repeat_interleave(W, 0).contiguous()
            lens_dev = _i32_dev([int(l) for l in enc_len for _ in range(W)], dev)
            z = lambda *s: torch.zeros(*s, device=dev)
            h1, c1, h2, c2 = z(NW, Sd), z(NW, Sd), z(NW, Sd), z(NW, Sd)
            score = torch.full((N, W), float('-inf'), device=dev)
            score[:, 0] = 0.0                                                 # step 0 expands hypothesis 0 only
            fin = torch.zeros(N, W, dtype=torch.int32, device=dev)
            tok = torch.zeros(NW, dtype=torch.int32, device=dev)              # <SOS> = 0 (asr.py:134)
            last = self.embed.weight[0:1].expand(NW, Sd).contiguous()
            use_lm = rnn_lm is not None and lm_weight != 0
            if use_lm:
                g1, g2 = rnn_lm.init_hidden(NW, dev)
            parents = torch.zeros(max_steps, N, W, dtype=torch.int32, device=dev)
            tokens = torch.zeros(max_steps, N, W, dtype=torch.int32, device=dev)
            wc, bc = self.char_trans.weight.contiguous(), self.char_trans.bias.contiguous()
            emb = self.embed.weight.contiguous()
            check_every = int(self.decode_stop_check or 0)
            steps = 0

            def gather(src, idx, group):
                dst = torch.empty(idx.numel(), src.shape[1], device=dev, dtype=src.dtype)
                _lib.check(lib.ssasr_gather_rows(src.data_ptr(), dst.data_ptr(), idx.data_ptr(), idx.numel(),
                                                 src.shape[1] * src.element_size(), group, st), 'ssasr_gather_rows')
                return dst
            for t in range(max_steps):
                _, ctx = Fk.attn_step(h1, enc_r, psi_r, lens_dev, self.attention.phi.weight)
                h1n, c1n = Fk.lstm_cell(torch.cat([last, ctx], -1), h1, c1, self.decoder.layer_1)
                h2n, c2n = Fk.lstm_cell(h1n, h2, c2, self.decoder.layer_2)
                logits = torch.empty(NW, C, device=dev)
                _lib.check(lib.ssasr_gemm_f32(NW, C, Sd, h2n.data_ptr(), Sd, 1, wc.data_ptr(), Sd, 1, logits.data_ptr(), C,
                                              bc.data_ptr(), 0, 0, st), 'ssasr_gemm_f32')
                lm_logits = None
                if use_lm:
                    lm_logits, (g1n, g2n) = rnn_lm(tok, g1, g2)
                    lm_logits = lm_logits.float().contiguous()
                score_n, fin_n = torch.empty_like(score), torch.empty_like(fin)
                _lib.check(lib.ssasr_beam_select(logits.data_ptr(), lm_logits.data_ptr() if use_lm else None, float(lm_weight), N, W,
                                                 C, int(eos_id), score.data_ptr(), fin.data_ptr(), score_n.data_ptr(),
                                                 fin_n.data_ptr(), parents[t].data_ptr(), tokens[t].data_ptr(), st),
                           'ssasr_beam_select')
                par = parents[t].view(-1)
                h1, c1, h2, c2 = (gather(v, par, W) for v in (h1n, c1n, h2n, c2n))
                if use_lm:
                    g1, g2 = gather(g1n.contiguous(), par, W), gather(g2n.contiguous(), par, W)
                tok = tokens[t].view(-1)
                last = gather(emb, tok, 0)
                score, fin = score_n, fin_n
                steps = t + 1
                if check_every > 0 and steps % check_every == 0 and steps < max_steps and bool(fin.all()):
                    break
        self.last_decode_steps = steps
        # back-track the best survivor of every utterance on the host
        par = parents[:steps].cpu().numpy()
        tk = tokens[:steps].cpu().numpy()
        sc = score.cpu().numpy()
        out, best_scores = [], []
        for n in range(N):
            w = int(np.argmax(sc[n]))                       # first maximum: ties -> lower index
            best_scores.append(float(sc[n, w]))
            ids = []
            for t in range(steps - 1, -1, -1):
                ids.append(int(tk[t, n, w]))
                w = int(par[t, n, w])
            ids.reverse()
            if eos_id in ids:
                ids = ids[:ids.index(eos_id)]
            out.append(ids)
        return (out, best_scores) if return_scores else out

    def beam_decode(self, x, x_len, rnn_lm, mapper, lm_weight, beam_size):
        """ASR.decode's signature (asr.py:112) plus the beam size the reference configures but never uses (trainer.py:554,590)."""
        assert len(x.shape) == 3 and x.shape[0] == 1
        ids = self.beam_decode_batch(x, x_len, beam_size, rnn_lm=rnn_lm, lm_weight=lm_weight,
                                     eos_id=int(mapper.char_to_ind(EOS_TKN)))[0]
        return ''.join(mapper.ind_to_char(i) for i in ids)

    # ------------------------------------------------------------------------------------------
    def init_parameters(self):
        """asr.py:175-212."""
        for p in self.parameters():
            data = p.data
            if data.dim() == 1:
                data.zero_()
            elif data.dim() == 2:
                data.normal_(0, 1. / math.sqrt(data.size(1)))
            else:
                raise NotImplementedError
        self.embed.weight.data.normal_(0, 1)
        for bias in (self.decoder.layer_1.bias_ih, self.decoder.layer_2.bias_ih):
            n = bias.size(0)
            bias.data[n // 4:n // 2].fill_(1.)
