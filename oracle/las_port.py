"""CPU "port" of the reference's LAS train / decode step built from the SAME torch library
calls the reference dispatches (cuDNN-style nn.LSTM on packed sequences, nn.LSTMCell,
bmm attention, CrossEntropyLoss, clip_grad_norm_, Adadelta).

TEST / BASELINE INFRASTRUCTURE ONLY: used by `bench.py` for the `cpu_baseline` object and the
`--impl reference` arm (the reference itself is Python and does not travel to the GPU box), and
by tests/ to confirm it agrees with oracle/las_oracle.py and the committed golden vectors.  It
exists so that the CPU figure reported beside the B200 numbers has the reference's own cost
profile (per-timestep packed LSTM, autograd through slices), not the cost of the explicit-loop
oracle.  Call sequence mirrored: /root/reference/src/asr.py:52-110 (forward), :112-173
(decode), /root/reference/src/trainer.py:415-438 + :131-148 (loss, backward, clip, step).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as Fn
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

from . import las_oracle as O


class Port:
    def __init__(self, sd, tf_rate=1.0):
        g = lambda k: sd[k]
        S = g('encoder.blstm_4.weight_hh_l0').shape[1]
        F = g('encoder.blstm_1.layer.weight_ih_l0').shape[1]
        Sd = g('decoder.layer_1.weight_hh').shape[1]
        M = g('attention.phi.weight').shape[0]
        C = g('char_trans.weight').shape[0]
        self.m = nn.ModuleDict({
            'encoder_blstm_1_layer': nn.LSTM(F, S, bidirectional=True, batch_first=True),
            'encoder_blstm_2_layer': nn.LSTM(4 * S, S, bidirectional=True, batch_first=True),
            'encoder_blstm_3_layer': nn.LSTM(4 * S, S, bidirectional=True, batch_first=True),
            'encoder_blstm_4': nn.LSTM(4 * S, S, bidirectional=True),
            'attention_phi': nn.Linear(Sd, M, bias=False),
            'attention_psi': nn.Linear(2 * S, M),
            'decoder_layer_1': nn.LSTMCell(2 * S + Sd, Sd),
            'decoder_layer_2': nn.LSTMCell(Sd, Sd),
            'embed': nn.Embedding(C, Sd),
            'char_trans': nn.Linear(Sd, C),
        })
        own = dict(self.m.named_parameters())
        with torch.no_grad():
            for k, v in sd.items():
                mod, _, par = k.rpartition('.')
                own[mod.replace('.', '_') + '.' + par].copy_(v)
        self.tf_rate = tf_rate
        self.Sd = Sd

    def parameters(self):
        return self.m.parameters()

    def grads(self):
        out = {}
        for k, p in self.m.named_parameters():
            mod, _, par = k.rpartition('.')
            name = {'encoder_blstm_1_layer': 'encoder.blstm_1.layer', 'encoder_blstm_2_layer': 'encoder.blstm_2.layer',
                    'encoder_blstm_3_layer': 'encoder.blstm_3.layer', 'encoder_blstm_4': 'encoder.blstm_4',
                    'attention_phi': 'attention.phi', 'attention_psi': 'attention.psi',
                    'decoder_layer_1': 'decoder.layer_1', 'decoder_layer_2': 'decoder.layer_2',
                    'embed': 'embed', 'char_trans': 'char_trans'}[mod]
            out[name + '.' + par] = p.grad
        return out

    # -- encoder -------------------------------------------------------------------------------
    def listen(self, x, lens):
        for k in ('encoder_blstm_1_layer', 'encoder_blstm_2_layer', 'encoder_blstm_3_layer'):
            packed = pack_padded_sequence(x, lens, batch_first=True)
            out, _ = self.m[k](packed)
            out, l2 = pad_packed_sequence(out, batch_first=True)
            B, T, D = out.shape
            x = out[:, :T - (T % 2)].contiguous().view(B, T // 2, 2 * D)
            lens = [int(s / 2) for s in l2.tolist()]
        x, _ = self.m['encoder_blstm_4'](x)
        return x, lens

    # -- decoder -------------------------------------------------------------------------------
    def _attend(self, s, psi, mask, enc):
        q = torch.tanh(self.m['attention_phi'](s))
        e = torch.bmm(psi, q.unsqueeze(2)).squeeze(2)
        e = e.masked_fill(mask, float('-inf'))
        a = torch.softmax(e, -1)
        return a, torch.bmm(a.unsqueeze(1), enc).squeeze(1)

    def forward(self, x, decode_step, teacher=None, lens=None, rng=None):
        enc, el = self.listen(x, lens)
        B = x.shape[0]
        mask = torch.arange(enc.shape[1])[None, :] >= torch.tensor(el)[:, None]
        psi = torch.tanh(self.m['attention_psi'](enc))
        temb = self.m['embed'](teacher) if teacher is not None else None
        h1 = c1 = h2 = c2 = enc.new_zeros(B, self.Sd)
        last = self.m['embed'](torch.zeros(B, dtype=torch.long))
        logits, atts = [], []
        for t in range(decode_step):
            a, ctx = self._attend(h1, psi, mask, enc)
            h1, c1 = self.m['decoder_layer_1'](torch.cat([last, ctx], -1), (h1, c1))
            h2, c2 = self.m['decoder_layer_2'](h1, (h2, c2))
            cur = self.m['char_trans'](h2)
            if temb is not None:
                draw = rng.random() if rng is not None else 0.0
                if draw <= self.tf_rate:
                    last = temb[:, t + 1, :]
                else:
                    last = self.m['embed'](torch.distributions.Categorical(Fn.softmax(cur, -1)).sample())
            else:
                last = self.m['embed'](torch.argmax(cur, -1))
            logits.append(cur)
            atts.append(a.detach())
        return el, torch.stack(logits, 1), torch.stack(atts, 1)

    def train_step(self, x, lens, y, optim=None, grad_clip=5.0, rng=None):
        """trainer.py:415-438 + Solver.step :131-148. Returns (loss, logits)."""
        y_lens = [int(l) + 1 for l in torch.sum(y != 0, dim=-1)]
        ans_len = max(y_lens) - 1
        if optim is not None:
            optim.zero_grad()
        else:
            for p in self.parameters():
                p.grad = None
        _, pred, _ = self.forward(x, ans_len, teacher=y, lens=lens, rng=rng)
        label = y[:, 1:ans_len + 1].contiguous()
        b, t, c = pred.shape
        loss = Fn.cross_entropy(pred.view(b * t, c), label.view(-1), ignore_index=0, reduction='none')
        loss = torch.sum(loss.view(b, t), dim=-1) / torch.sum(y != 0, dim=-1).to(torch.float32)
        loss = torch.mean(loss)
        loss.backward()
        if optim is not None:
            gn = nn.utils.clip_grad_norm_(self.parameters(), grad_clip)
            if not math.isnan(float(gn)):
                optim.step()
        return loss.detach(), pred.detach()

    @torch.no_grad()
    def decode(self, x, x_len, lm=None, lm_weight=0.0, max_steps=O.MAX_DECODE):
        """asr.py:112-173 for one utterance; returns emitted token ids."""
        enc, el = self.listen(x, x_len)
        mask = torch.arange(enc.shape[1])[None, :] >= torch.tensor(el)[:, None]
        psi = torch.tanh(self.m['attention_psi'](enc))
        h1 = c1 = h2 = c2 = enc.new_zeros(1, self.Sd)
        last_idx = torch.zeros(1, dtype=torch.long)
        last = self.m['embed'](last_idx)
        if lm is not None:
            H = lm['layer_1.weight_hh'].shape[1]
            g1 = g2 = enc.new_zeros(1, H)
        out = []
        while len(out) < max_steps:
            a, ctx = self._attend(h1, psi, mask, enc)
            h1, c1 = self.m['decoder_layer_1'](torch.cat([last, ctx], -1), (h1, c1))
            h2, c2 = self.m['decoder_layer_2'](h1, (h2, c2))
            final = Fn.log_softmax(self.m['char_trans'](h2), -1)
            if lm is not None:
                lo, g1, g2 = O.charlm_step(lm, last_idx, g1, g2)
                final = final + lm_weight * Fn.log_softmax(lo, -1)
            last_idx = torch.argmax(final, -1)
            last = self.m['embed'](last_idx)
            if int(last_idx) == O.EOS_ID:
                break
            out.append(int(last_idx))
        return out
