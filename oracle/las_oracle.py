"""CPU oracle for the Listen-Attend-Spell hot path of cadia-lvl/ss_asr.

TEST INFRASTRUCTURE ONLY.  Nothing in `ss_asr_b200/` may import this module; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg do, and there only as the checker / the CPU arm.

This is an explicit-loop restatement (plain torch tensor arithmetic on the CPU, any
float dtype) of what the reference computes, written from the reference's semantics:

  * Listener / pBLSTM  ........ /root/reference/src/asr.py:214-264, 394-450
  * Attention ................. /root/reference/src/asr.py:328-392
  * Speller ................... /root/reference/src/asr.py:267-326
  * ASR.forward / ASR.decode .. /root/reference/src/asr.py:52-173
  * ASR.init_parameters ....... /root/reference/src/asr.py:175-212
  * ASR loss .................. /root/reference/src/trainer.py:394-395,426-434
  * CharLM .................... /root/reference/src/charlm.py:46-61

Parity pinning: the reference holds no golden vectors, so the oracle
is pinned against outputs of the UNMODIFIED reference run through
oracle/ref_shim.py in the authoring container; those outputs are committed
under tests/golden/*.npz with the script that made them (tests/golden/make_golden.py)
and re-checked by tests/test_oracle_golden.py on every run.

All functions take a flat `sd` (state_dict-like mapping name -> tensor) using the
reference's parameter names.
"""
import math

import torch

EOS_ID = 1          # '>' ; preprocess.py:24-27, ASRDataset.py:233-235
SOS_ID = 0          # '<' ; also the padding id
MAX_DECODE = 200    # asr.py:128
TOKENS = '<>$' + 'abcdefghijklmnoprstuvxy0123456789' + 'áéíóúýæöþð' + ' .,?'  # preprocess.py:17-27


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------
def make_state_dict(output_dim=50, encoder_state_size=256, decoder_state_size=256,
                    mlp_out_size=128, feature_dim=80, seed=1, dtype=torch.float32):
    """Seeded parameters drawn the way ASR.__init__ + init_parameters draw them
    (asr.py:31-50, 175-212): torch.nn containers are constructed in the reference's
    order (so the global RNG is consumed identically), then every 1-D parameter is
    zeroed, every 2-D parameter ~ N(0, 1/fan_in), embed ~ N(0,1), and the forget-gate
    quarter of decoder.layer_{1,2}.bias_ih is set to 1."""
    import torch.nn as nn
    torch.manual_seed(seed)
    S, Sd, F, M = encoder_state_size, decoder_state_size, feature_dim, mlp_out_size
    mods = [
        ('encoder.blstm_1.layer', nn.LSTM(F, S, bidirectional=True, batch_first=True)),
        ('encoder.blstm_2.layer', nn.LSTM(4 * S, S, bidirectional=True, batch_first=True)),
        ('encoder.blstm_3.layer', nn.LSTM(4 * S, S, bidirectional=True, batch_first=True)),
        ('encoder.blstm_4', nn.LSTM(4 * S, S, bidirectional=True)),
        ('attention.phi', nn.Linear(Sd, M, bias=False)),
        ('attention.psi', nn.Linear(2 * S, M)),
        ('decoder.layer_1', nn.LSTMCell(2 * S + Sd, Sd)),
        ('decoder.layer_2', nn.LSTMCell(Sd, Sd)),
        ('embed', nn.Embedding(output_dim, Sd)),
        ('char_trans', nn.Linear(Sd, output_dim)),
    ]
    sd = {}
    for prefix, m in mods:
        for n, p in m.named_parameters():
            sd[prefix + '.' + n] = p.data
    for n, p in sd.items():
        if p.dim() == 1:
            p.zero_()
        else:
            p.normal_(0, 1.0 / math.sqrt(p.size(1)))
    sd['embed.weight'].normal_(0, 1)
    for l in ('decoder.layer_1.bias_ih', 'decoder.layer_2.bias_ih'):
        n = sd[l].numel()
        sd[l][n // 4:n // 2] = 1.0
    return {k: v.clone().to(dtype) for k, v in sd.items()}


def make_charlm_state_dict(input_size=50, hidden_size=128, seed=7, dtype=torch.float32):
    """CharLM(input_size, hidden_size) default torch init (charlm.py:5-44); ASRTester uses a
    freshly initialised, never-loaded LM (trainer.py:567-569)."""
    import torch.nn as nn
    torch.manual_seed(seed)
    mods = [('emb', nn.Embedding(input_size, hidden_size)),
            ('layer_1', nn.GRUCell(hidden_size, hidden_size)),
            ('layer_2', nn.GRUCell(hidden_size, hidden_size)),
            ('out', nn.Linear(hidden_size, input_size))]
    sd = {}
    for prefix, m in mods:
        for n, p in m.named_parameters():
            sd[prefix + '.' + n] = p.data.clone().to(dtype)
    return sd


# ----------------------------------------------------------------------------------------------
# synthetic inputs (SURVEY.md §8d "common synthetic recipe")
# ----------------------------------------------------------------------------------------------
def synth_batch(B, T, F, U, seed=1234, n_tokens=50):
    g = torch.Generator('cpu').manual_seed(seed)
    lens = torch.randint(3 * T // 4, T + 1, (B,), generator=g)
    lens, _ = torch.sort(lens, descending=True)
    lens[0] = T
    x = torch.randn(B, T, F, generator=g)
    t_idx = torch.arange(T)[None, :, None]
    x = x * (t_idx < lens[:, None, None]).to(x.dtype)
    ylen = torch.randint(max(1, U // 2), U + 1, (B,), generator=g)
    y = torch.zeros(B, U + 2, dtype=torch.long)
    tok = torch.randint(3, n_tokens, (B, U + 2), generator=g)
    for i in range(B):
        n = int(ylen[i])
        y[i, 1:1 + n] = tok[i, 1:1 + n]
        y[i, 1 + n] = EOS_ID
    if int(ylen.max()) < U:        # make ans_len == U + 1 deterministic
        y[0, 1:1 + U] = tok[0, 1:1 + U]
        y[0, 1 + U] = EOS_ID
    return x, [int(v) for v in lens], y


# ----------------------------------------------------------------------------------------------
# LSTM pieces
# ----------------------------------------------------------------------------------------------
def _cell(gates, c):
    """PyTorch LSTM cell, gate order i,f,g,o (Appendix A of SURVEY.md)."""
    S = c.shape[-1]
    i = torch.sigmoid(gates[..., 0:S])
    f = torch.sigmoid(gates[..., S:2 * S])
    g = torch.tanh(gates[..., 2 * S:3 * S])
    o = torch.sigmoid(gates[..., 3 * S:4 * S])
    c2 = f * c + i * g
    return o * torch.tanh(c2), c2


def _lstm_dir_params(sd, prefix, reverse):
    sfx = '_l0_reverse' if reverse else '_l0'
    return (sd[prefix + 'weight_ih' + sfx], sd[prefix + 'weight_hh' + sfx],
            sd[prefix + 'bias_ih' + sfx], sd[prefix + 'bias_hh' + sfx])


def blstm_packed(sd, prefix, x, lens):
    """Bidirectional LSTM with pack_padded_sequence semantics (asr.py:410-418): for utterance i
    the forward direction runs t=0..len-1, the reverse direction t=len-1..0, both from zero
    state; outputs beyond len are exactly zero; output time dim = max(lens)."""
    B = x.shape[0]
    Tm = max(lens)
    lens_t = torch.tensor(lens)
    outs = []
    for reverse in (False, True):
        Wih, Whh, bih, bhh = _lstm_dir_params(sd, prefix, reverse)
        S = Whh.shape[1]
        xp = x[:, :Tm] @ Wih.t() + (bih + bhh)
        h = x.new_zeros(B, S)
        c = x.new_zeros(B, S)
        out = [None] * Tm
        order = range(Tm - 1, -1, -1) if reverse else range(Tm)
        for t in order:
            valid = (t < lens_t)[:, None].to(x.dtype)
            h2, c2 = _cell(xp[:, t] + h @ Whh.t(), c)
            h = valid * h2 + (1 - valid) * h * (0.0 if reverse else 1.0)
            c = valid * c2 + (1 - valid) * c * (0.0 if reverse else 1.0)
            out[t] = valid * h2
        outs.append(torch.stack(out, 1))
    return torch.cat(outs, -1)


def downsample(x):
    """asr.py:429-450: drop an odd last frame, concat frame pairs on the feature axis."""
    B, T, Fd = x.shape
    T2 = T // 2
    return x[:, :2 * T2].reshape(B, T2, 2 * Fd)


def pblstm(sd, prefix, x, lens):
    out = downsample(blstm_packed(sd, prefix, x, lens))
    return out, [int(s / 2) for s in lens]


def blstm_seqfirst(sd, prefix, x):
    """encoder.blstm_4 (asr.py:237-238,262): nn.LSTM WITHOUT batch_first fed a [B,T',4S] tensor:
    dim 0 (utterances) is the time axis, dim 1 (frames) the batch; unpacked, zero init."""
    L, N = x.shape[0], x.shape[1]
    outs = []
    for reverse in (False, True):
        Wih, Whh, bih, bhh = _lstm_dir_params(sd, prefix, reverse)
        S = Whh.shape[1]
        xp = x @ Wih.t() + (bih + bhh)
        h = x.new_zeros(N, S)
        c = x.new_zeros(N, S)
        out = [None] * L
        for s in (range(L - 1, -1, -1) if reverse else range(L)):
            h, c = _cell(xp[s] + h @ Whh.t(), c)
            out[s] = h
        outs.append(torch.stack(out, 0))
    return torch.cat(outs, -1)


def listener(sd, x, lens, prefix='encoder.'):
    """asr.py:243-264."""
    x, lens = pblstm(sd, prefix + 'blstm_1.layer.', x, lens)
    x, lens = pblstm(sd, prefix + 'blstm_2.layer.', x, lens)
    x, lens = pblstm(sd, prefix + 'blstm_3.layer.', x, lens)
    x = blstm_seqfirst(sd, prefix + 'blstm_4.', x)
    return x, lens


# ----------------------------------------------------------------------------------------------
# attention / speller
# ----------------------------------------------------------------------------------------------
def attention_memory(sd, enc, enc_lens):
    """asr.py:363-381: pad mask (True where j >= len) and psi~ = tanh(psi(enc))."""
    B, Tp, _ = enc.shape
    mask = torch.arange(Tp)[None, :] >= torch.tensor(enc_lens)[:, None]
    psi = torch.tanh(enc @ sd['attention.psi.weight'].t() + sd['attention.psi.bias'])
    return psi, mask


def attention_step(sd, s, psi, mask, enc):
    """asr.py:383-392."""
    q = torch.tanh(s @ sd['attention.phi.weight'].t())
    e = (psi * q[:, None, :]).sum(-1)
    e = e.masked_fill(mask, float('-inf'))
    a = torch.softmax(e, -1)
    ctx = (a[:, :, None] * enc).sum(1)
    return a, ctx


def speller_step(sd, inp, state):
    """asr.py:314-326: two stacked LSTMCells."""
    (h1, c1), (h2, c2) = state
    p = 'decoder.layer_1.'
    g1 = inp @ sd[p + 'weight_ih'].t() + sd[p + 'bias_ih'] + h1 @ sd[p + 'weight_hh'].t() + sd[p + 'bias_hh']
    h1, c1 = _cell(g1, c1)
    p = 'decoder.layer_2.'
    g2 = h1 @ sd[p + 'weight_ih'].t() + sd[p + 'bias_ih'] + h2 @ sd[p + 'weight_hh'].t() + sd[p + 'bias_hh']
    h2, c2 = _cell(g2, c2)
    return h2, ((h1, c1), (h2, c2))


def spell(sd, enc, enc_lens, decode_step, teacher=None, tf_mask=None, sampled=None):
    """Decoder loop of ASR.forward (asr.py:65-110).

    teacher: LongTensor [B, L] or None (greedy).  tf_mask[t] False => the next input is
    embed(sampled[:, t]) (the reference samples from Categorical(softmax); the oracle
    replays externally supplied samples).  Returns logits [B,U,C], att [B,U,T']."""
    B = enc.shape[0]
    Sd = sd['decoder.layer_1.weight_hh'].shape[1]
    emb = sd['embed.weight']
    psi, mask = attention_memory(sd, enc, enc_lens)
    z = enc.new_zeros(B, Sd)
    state = ((z, z), (z, z))
    last = emb[torch.zeros(B, dtype=torch.long)]
    logits, atts, toks = [], [], []
    for t in range(decode_step):
        a, ctx = attention_step(sd, state[0][0], psi, mask, enc)
        h2, state = speller_step(sd, torch.cat([last, ctx], -1), state)
        cur = h2 @ sd['char_trans.weight'].t() + sd['char_trans.bias']
        if teacher is not None:
            if tf_mask is None or tf_mask[t]:
                nxt = teacher[:, t + 1]
            else:
                nxt = sampled[:, t]
        else:
            nxt = torch.argmax(cur, -1)
        last = emb[nxt]
        toks.append(nxt)
        logits.append(cur)
        atts.append(a)
    return torch.stack(logits, 1), torch.stack(atts, 1), torch.stack(toks, 1)


def asr_forward(sd, x, lens, decode_step, teacher=None, tf_mask=None, sampled=None):
    """ASR.forward (asr.py:52-110) -> (encode_len, logits [B,U,C], att [B,U,T'], enc)."""
    enc, enc_lens = listener(sd, x, lens)
    logits, att, _ = spell(sd, enc, enc_lens, decode_step, teacher, tf_mask, sampled)
    return enc_lens, logits, att, enc


def asr_loss(logits, y):
    """trainer.py:426-434: CE(ignore_index=0,'none') summed per utterance, divided by
    count(y != 0), mean over the batch."""
    B, U, C = logits.shape
    label = y[:, 1:U + 1]
    lp = torch.log_softmax(logits, -1)
    nll = -lp.gather(-1, label[..., None]).squeeze(-1)
    nll = nll * (label != 0).to(logits.dtype)
    return (nll.sum(-1) / (y != 0).sum(-1).to(logits.dtype)).mean()


def train_step_grads(sd, x, lens, y, dtype=torch.float32):
    """Loss and gradients for one teacher-forced (tf_rate=1) batch, trainer.py:415-437."""
    p = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
    U = int(max((y != 0).sum(-1) + 1)) - 1          # ans_len, trainer.py:418 + ASRDataset.py:338
    _, logits, att, enc = asr_forward(p, x.to(dtype), lens, U, teacher=y)
    loss = asr_loss(logits, y)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    return loss.detach(), logits.detach(), att.detach(), enc.detach(), grads


# ----------------------------------------------------------------------------------------------
# CharLM + greedy decode
# ----------------------------------------------------------------------------------------------
def _gru_cell(sd, p, x, h):
    """PyTorch GRUCell, gate order r,z,n."""
    H = h.shape[-1]
    gi = x @ sd[p + 'weight_ih'].t() + sd[p + 'bias_ih']
    gh = h @ sd[p + 'weight_hh'].t() + sd[p + 'bias_hh']
    r = torch.sigmoid(gi[..., :H] + gh[..., :H])
    z = torch.sigmoid(gi[..., H:2 * H] + gh[..., H:2 * H])
    n = torch.tanh(gi[..., 2 * H:] + r * gh[..., 2 * H:])
    return (1 - z) * n + z * h


def charlm_step(lm, idx, h1, h2):
    """charlm.py:46-57."""
    x = lm['emb.weight'][idx]
    h1 = _gru_cell(lm, 'layer_1.', x, h1)
    h2 = _gru_cell(lm, 'layer_2.', h1, h2)
    return h2 @ lm['out.weight'].t() + lm['out.bias'], h1, h2


def decode_greedy(sd, x, x_len, lm=None, lm_weight=0.0, max_steps=MAX_DECODE, return_margin=False):
    """ASR.decode (asr.py:112-173) for ONE utterance x [1,T,F]; returns the emitted token ids
    (EOS not included).  With lm=None the LM term is dropped (identical to lm_weight=0 as far
    as the argmax is concerned)."""
    assert x.dim() == 3 and x.shape[0] == 1
    enc, enc_lens = listener(sd, x, x_len)
    psi, mask = attention_memory(sd, enc, enc_lens)
    Sd = sd['decoder.layer_1.weight_hh'].shape[1]
    z = enc.new_zeros(1, Sd)
    state = ((z, z), (z, z))
    emb = sd['embed.weight']
    last = emb[torch.zeros(1, dtype=torch.long)]
    last_idx = torch.zeros(1, dtype=torch.long)
    if lm is not None:
        H = lm['layer_1.weight_hh'].shape[1]
        g1 = enc.new_zeros(1, H)
        g2 = enc.new_zeros(1, H)
    out, margin = [], float('inf')
    while len(out) < max_steps:
        a, ctx = attention_step(sd, state[0][0], psi, mask, enc)
        h2, state = speller_step(sd, torch.cat([last, ctx], -1), state)
        final = torch.log_softmax(h2 @ sd['char_trans.weight'].t() + sd['char_trans.bias'], -1)
        if lm is not None:
            lo, g1, g2 = charlm_step(lm, last_idx, g1, g2)
            final = final + lm_weight * torch.log_softmax(lo, -1)
        top2 = torch.topk(final[0], 2).values
        margin = min(margin, float(top2[0] - top2[1]))
        nxt = torch.argmax(final, -1)
        last_idx = nxt
        last = emb[nxt]
        if int(nxt) == EOS_ID:
            break
        out.append(int(nxt))
    return (out, margin) if return_margin else out


def decode_beam(sd, x, x_len, beam_size, lm=None, lm_weight=0.0, max_steps=MAX_DECODE, return_score=False):
    """Beam search over the step of ASR.decode (asr.py:143-172) for ONE utterance x [1,T,F].  NOT in the reference: it configures
    a beam (conf/default.yaml:16-19, trainer.py:552-554) and then decodes greedily (trainer.py:590 TODO), so these semantics are
    this repo's (stated here, implemented by ss_asr_b200.asr.ASR.beam_decode_batch + csrc/beam.cu) and are anchored on the
    reference by the identity  decode_beam(beam_size=1) == decode_greedy  (same step, same `final` score, same EOS rule):
      * score of a hypothesis = sum over its tokens (EOS included) of final = log_softmax(asr) [+ lm_weight * log_softmax(lm)];
      * each step every live hypothesis is extended by all C tokens, a finished one (EOS emitted) stays ONE candidate with its
        score; the `beam_size` best candidates survive, ties -> lower parent index, then lower token id;
      * stops when every survivor is finished or after max_steps tokens; returns the tokens (EOS excluded) of the best survivor
        (ties -> first)."""
    assert x.dim() == 3 and x.shape[0] == 1
    W = int(beam_size)
    enc, enc_lens = listener(sd, x, x_len)
    psi, mask = attention_memory(sd, enc, enc_lens)
    Sd = sd['decoder.layer_1.weight_hh'].shape[1]
    emb = sd['embed.weight']
    C = sd['char_trans.weight'].shape[0]
    z = enc.new_zeros(1, Sd)
    H = lm['layer_1.weight_hh'].shape[1] if lm is not None else 1
    g0 = enc.new_zeros(1, H)
    # hypothesis: (score, finished, tokens, last_idx, speller state, lm state)
    hyps = [(0.0, False, [], 0, ((z, z), (z, z)), (g0, g0))]
    for _ in range(max_steps):
        cands = []              # (score, parent, token, new state, new lm state)
        for w, (score, fin, toks, last_idx, state, (g1, g2)) in enumerate(hyps):
            if fin:
                cands.append((score, w, EOS_ID, state, (g1, g2)))
                continue
            _, ctx = attention_step(sd, state[0][0], psi, mask, enc)
            h2, nstate = speller_step(sd, torch.cat([emb[torch.tensor([last_idx])], ctx], -1), state)
            final = torch.log_softmax(h2 @ sd['char_trans.weight'].t() + sd['char_trans.bias'], -1)
            ng = (g1, g2)
            if lm is not None:
                lo, n1, n2 = charlm_step(lm, torch.tensor([last_idx]), g1, g2)
                final = final + lm_weight * torch.log_softmax(lo, -1)
                ng = (n1, n2)
            sc = (torch.tensor(score, dtype=final.dtype) + final[0])
            for c in range(C):
                cands.append((float(sc[c]), w, c, nstate, ng))
        cands.sort(key=lambda t: (-t[0], t[1], t[2]))
        new = []
        for score, w, c, nstate, ng in cands[:W]:
            _, fin, toks, _, _, _ = hyps[w]
            if fin:
                new.append(hyps[w])
            else:
                new.append((score, c == EOS_ID, toks + ([] if c == EOS_ID else [c]), c, nstate, ng))
        hyps = new
        if all(h[1] for h in hyps):
            break
    best = max(range(len(hyps)), key=lambda i: (hyps[i][0], -i))
    return (hyps[best][2], hyps[best][0]) if return_score else hyps[best][2]


def ids_to_str(ids):
    return ''.join(TOKENS[i] for i in ids)
