"""torch.autograd glue over the C ABI (include/ssasr.h).  PyTorch is used here for device memory, streams
and the autograd graph only; every FLOP of the hot path is in libssasr.so."""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import check, ptr, stream


_TC_RECURRENCE = os.environ.get('SSASR_TC_RECURRENCE', '1') != '0'   # bf16 path: recurrence on tcgen05
# bf16 path: the Speller's layer-2 cell chain (forward and backward) on a second stream inside the C call -- it never feeds
# the attention query (asr.py:84), so it leaves the dependent chain of a decoding step
_DUAL_STREAM_SPELLER = os.environ.get('SSASR_DUAL_STREAM_SPELLER', '1') != '0'
# bf16 path: teacher-forced runs of the attend-and-spell loop in ONE cluster-persistent kernel launch per run (spell_cl.cu)
_CLUSTER_SPELLER = os.environ.get('SSASR_SPELL_CL', '1') != '0'


_CLUSTER_SPELLER_BWD = os.environ.get('SSASR_SPELL_CL_BWD', '1') != '0'


def set_cluster_speller(on, backward=None):
    """on: forward (and, unless `backward` says otherwise, backward) of the decoder loop in the cluster-persistent kernels"""
    global _CLUSTER_SPELLER, _CLUSTER_SPELLER_BWD
    _CLUSTER_SPELLER = bool(on)
    _CLUSTER_SPELLER_BWD = bool(on if backward is None else backward)


def set_dual_stream_speller(on):
    global _DUAL_STREAM_SPELLER
    _DUAL_STREAM_SPELLER = bool(on)


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


def _consume(ctx, what):
    """Single-backward contract.  The backward kernels work IN PLACE on the tensors saved by the forward pass (the gate
    activations become the gate gradients, the bf16 h copy is masked, ...) through raw pointers autograd's version counters
    never see: a second backward over the same graph (`retain_graph=True`) would silently return wrong gradients, so it
    raises instead."""
    if getattr(ctx, 'consumed', False):
        raise RuntimeError('ss_asr_b200 %s: backward called a second time over the same forward pass; the CUDA backward kernels '
                           'overwrite the saved activations in place, so retain_graph=True is not supported -- run the '
                           'forward pass again' % what)
    ctx.consumed = True


def _join_after_backward():
    """Deferred weight gradients are joined when the autograd engine finishes the backward pass that produced them, so that
    ANY reader of `param.grad` after `loss.backward()` returns (torch's clip_grad_norm_ / optimisers as in the unmodified
    Solver.step, trainer.py:144-148) is ordered after the side stream."""
    if not _OVERLAP.get('cb_queued'):
        _OVERLAP['cb_queued'] = True

        def _cb():
            _OVERLAP['cb_queued'] = False
            join_deferred()
        torch.autograd.Variable._execution_engine.queue_callback(_cb)


# --------------------------------------------------------------------------------------------------
# bidirectional LSTM layer
# --------------------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------------------
# Deferred weight gradients.  The encoder's weight-gradient GEMMs (dW_ih, dW_hh) and their un-packing are not needed
# before the optimiser, while the next layer's recurrent backward kernel is latency-bound and occupies 64 of the 148
# SMs.  With `set_overlap_wgrad(True)` they are enqueued on a second stream (ordered after everything the layer's backward
# has launched) and `join_deferred()` makes the current stream wait for them.  OPT-IN.  The join happens automatically when
# the autograd engine finishes the backward pass (`_join_after_backward`), so every reader of `param.grad` after
# `loss.backward()` -- torch's own clip_grad_norm_ / optimisers included -- is ordered after the side stream; GradSync joins
# on the side stream itself (the bucket all-reduce is issued there).
# --------------------------------------------------------------------------------------------------
_OVERLAP = {'on': False, 'streams': {}, 'pending': [], 'cb_queued': False, 'deferred_total': 0}


def set_overlap_wgrad(on):
    join_deferred()
    _OVERLAP['on'] = bool(on)


def overlap_wgrad_enabled():
    return _OVERLAP['on']


def side_stream(device):
    device = torch.device(device)
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _OVERLAP['streams']:
        _OVERLAP['streams'][key] = torch.cuda.Stream(device)
    return _OVERLAP['streams'][key]


def join_deferred():
    """The current stream waits for every deferred weight-gradient computation; their workspaces are released."""
    if _OVERLAP['pending']:
        cur = torch.cuda.current_stream()
        for ev, _keep in _OVERLAP['pending']:
            cur.wait_event(ev)
        _OVERLAP['pending'] = []


# Side channel between consecutive encoder layers on the tensor-core path: the recurrent kernel of layer l also writes a bf16
# copy of its output (exchange buffer / weight-gradient operand).  Viewed [B, T/2, 4S] that copy IS the bf16 input of layer l+1,
# so the fp32 -> bf16 conversion of the layer input (268 MB read + 134 MB written at layer 2 of C4) is skipped when the caller
# hands it over with `hint_bf16_input` right before `blstm(...)`; `LAST_BLSTM['hb']` exposes the copy after the call.
_XBF_HINT = {'t': None}
LAST_BLSTM = {'hb': None}


def hint_bf16_input(t):
    _XBF_HINT['t'] = t


class _BLSTM(torch.autograd.Function):
    """One bidirectional LSTM layer over rows indexed (seq, batch) -- see ssasr_blstm_fwd_f32.

    time_major=True : x [B, T_h, K], seq axis = dim 1 with per-utterance `lens` (packed-sequence semantics,
                      asr.py:410-418).  T_h must be even and >= max(lens).
    time_major=False: x [L, N, K], seq axis = dim 0, no lengths (encoder.blstm_4 quirk, asr.py:262)."""

    @staticmethod
    def forward(ctx, x, lens_dev, time_major, precision, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r):
        lib = _lib.load()
        _lib.require_cuda(x, 'BLSTM')
        x_bf, _XBF_HINT['t'] = _XBF_HINT['t'], None
        LAST_BLSTM['hb'] = None
        x = _f32c(x)
        d0, d1, K = x.shape
        S = w_hh_f.shape[1]
        dev = x.device
        n_rows = d0 * d1
        st = stream()
        bf16 = precision == 'bf16'
        tc_rec = bf16 and S % 64 == 0 and S <= 512 and _TC_RECURRENCE
        ctx.param_refs = (w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r)
        ws = [_f32c(w) for w in (w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r)]
        bias_p = torch.empty(8 * S, device=dev)
        wih_p = whh_p = whhT_p = wih_bf = whh_bf = wihT_bf = whhT_bf = None
        Kp = (K + 7) // 8 * 8
        if tc_rec:
            # tensor-core path: every packed bf16 operand of the layer (forward and backward) in ONE launch
            bfe = lambda *sh: torch.empty(*sh, device=dev, dtype=torch.bfloat16)
            wih_bf, whh_bf, wihT_bf, whhT_bf = bfe(8 * S, Kp), bfe(8 * S, S), bfe(K, 8 * S), bfe(2 * S, 4 * S)
            check(lib.ssasr_pack_blstm_bf16(*[ptr(w) for w in ws], S, K, Kp, ptr(bias_p), ptr(wih_bf), ptr(whh_bf), ptr(wihT_bf),
                                            ptr(whhT_bf), st), 'ssasr_pack_blstm_bf16')
        else:
            wih_p = torch.empty(8 * S, K, device=dev)
            whh_p = torch.empty(2, 4 * S, S, device=dev)
            whhT_p = torch.empty(2, S, 4 * S, device=dev)
            check(lib.ssasr_pack_blstm(*[ptr(w) for w in ws], S, K, ptr(wih_p), ptr(bias_p), ptr(whh_p), ptr(whhT_p), st),
                  'ssasr_pack_blstm')
        xp = torch.empty(n_rows, 8 * S, device=dev)
        hout = torch.empty(d0, d1, 2 * S, device=dev)
        cbuf = torch.empty(d0, d1, 2 * S, device=dev)
        bar = torch.zeros(4096, dtype=torch.int32, device=dev)
        if time_major:
            n_seq, n_batch, rs_seq, rs_batch = d1, d0, 1, d1
        else:
            n_seq, n_batch, rs_seq, rs_batch = d0, d1, d1, 1
        if bf16:
            given = (tc_rec and x_bf is not None and Kp == K and x_bf.dtype == torch.bfloat16 and x_bf.is_cuda and
                     x_bf.is_contiguous() and x_bf.numel() == n_rows * K)
            if given:
                xb = x_bf.view(n_rows, Kp)            # the previous layer's bf16 output: no conversion pass
            else:
                xb = torch.zeros(n_rows, Kp, device=dev, dtype=torch.bfloat16) if Kp != K else \
                    torch.empty(n_rows, Kp, device=dev, dtype=torch.bfloat16)
            hb = None
            if tc_rec:
                hb = torch.empty(n_rows, 2 * S, device=dev, dtype=torch.bfloat16)
                LAST_BLSTM['hb'] = hb
            else:
                wih_bf = torch.zeros(8 * S, Kp, device=dev, dtype=torch.bfloat16)
                check(lib.ssasr_cvt_bf16(ptr(wih_p), K, ptr(wih_bf), Kp, 8 * S, K, st), 'ssasr_cvt_bf16')
            check(lib.ssasr_blstm_fwd_bf16(None if given else ptr(x), n_rows, K, Kp, ptr(wih_bf), ptr(bias_p), ptr(whh_p), S, n_seq, n_batch,
                                           rs_seq, rs_batch, ptr(lens_dev) if time_major else None, ptr(xb), ptr(xp),
                                           ptr(hout), ptr(cbuf), ptr(bar), ptr(whh_bf), ptr(hb), st),
                  'ssasr_blstm_fwd_bf16')
        else:
            tws = x3 = None
            if precision == 'tf32x3' and not torch.is_grad_enabled():       # forward-only fast exact path
                if K % 4 == 0:
                    tws = torch.empty(2 * (n_rows + 8 * S) * K, device=dev)
                if S % 64 == 0 and S <= 256:
                    x3 = torch.empty(16 * S * S + 4 * n_rows * S, device=dev, dtype=torch.bfloat16)
                elif S == 512:
                    x3 = torch.empty(16 * S * S, device=dev, dtype=torch.bfloat16)      # 16-CTA cluster kernel: W_hh parts only
            check(lib.ssasr_blstm_fwd_f32(ptr(x), n_rows, K, ptr(wih_p), ptr(bias_p), ptr(whh_p), S, n_seq, n_batch, rs_seq,
                                          rs_batch, ptr(lens_dev) if time_major else None, ptr(xp), ptr(hout), ptr(cbuf),
                                          ptr(bar), ptr(tws), ptr(x3), st), 'ssasr_blstm_fwd_f32')
        ctx.bf16 = bf16
        ctx.fwd_bf = (xb, hb, Kp) if (bf16 and hb is not None) else None      # bf16 x / h copies reused by the weight gradients
        ctx.packed_bf = (wihT_bf, whhT_bf) if tc_rec else None                # backward operands packed with the forward ones
        e0 = torch.empty(0, device=dev)
        ctx.save_for_backward(x, wih_p if wih_p is not None else e0, whhT_p if whhT_p is not None else e0, xp, hout, cbuf,
                              lens_dev if time_major else torch.empty(0))
        ctx.geom = (n_rows, K, S, n_seq, n_batch, rs_seq, rs_batch, time_major, d1)
        ctx.need_dx = ctx.needs_input_grad[0]
        return hout

    @staticmethod
    def backward(ctx, dhout):
        lib = _lib.load()
        _consume(ctx, 'BLSTM')
        x, wih_p, whhT_p, act, hout, cbuf, lens_dev = ctx.saved_tensors
        n_rows, K, S, n_seq, n_batch, rs_seq, rs_batch, time_major, d1 = ctx.geom
        dev = x.device
        st = stream()
        dhout = _f32c(dhout)
        dx = torch.empty_like(x) if ctx.need_dx else None
        dwih_p = torch.empty(8 * S, K, device=dev)
        dbias_p = torch.empty(8 * S, device=dev)
        dwhh_p = torch.empty(2, 4 * S, S, device=dev)
        dcs = torch.empty(n_batch, 2 * S, device=dev)
        bar = torch.zeros(512, dtype=torch.int32, device=dev)
        if ctx.bf16:
            Rp = (n_rows + 7) // 8 * 8
            bf = lambda *s: torch.empty(*s, device=dev, dtype=torch.bfloat16)
            tc_rec = S % 64 == 0 and S <= 512 and _TC_RECURRENCE
            if ctx.packed_bf is not None:
                wihT_bf, whhT_bf = ctx.packed_bf
            else:
                wihT_bf = bf(K, 8 * S)
                check(lib.ssasr_cvt_bf16_t(ptr(wih_p), K, ptr(wihT_bf), 8 * S, 8 * S, K, 0, 0, 0, 0, st), 'ssasr_cvt_bf16_t')
                whhT_bf = None
                if tc_rec:
                    whhT_bf = bf(2 * S, 4 * S)
                    check(lib.ssasr_cvt_bf16(ptr(whhT_p), 4 * S, ptr(whhT_bf), 4 * S, 2 * S, 4 * S, st), 'ssasr_cvt_bf16')
            direct = tc_rec and ctx.fwd_bf is not None
            ws = [bf(n_rows, 8 * S) if (ctx.need_dx or tc_rec) else None] + \
                ([None, None, None] if direct else [bf(8 * S, Rp), bf(K, Rp), bf(2 * S, Rp)])
            xb_s, hb_s, Kp_s = ctx.fwd_bf if direct else (None, None, 0)
            # deferring is only safe when autograd will ADOPT the returned buffers (fresh .grad, i.e. zero_grad(set_to_none=True)):
            # accumulating into an existing .grad would read them on the main stream before the side stream has written them
            fresh = all(getattr(w, 'grad', None) is None for w in ctx.param_refs)
            side = side_stream(dev) if (direct and _OVERLAP['on'] and fresh) else None
            # gradient outputs are allocated BEFORE the call: anything still using that memory is ordered before the side stream
            g = [torch.empty(4 * S, K, device=dev), torch.empty(4 * S, S, device=dev), torch.empty(4 * S, device=dev),
                 torch.empty(4 * S, device=dev), torch.empty(4 * S, K, device=dev), torch.empty(4 * S, S, device=dev),
                 torch.empty(4 * S, device=dev), torch.empty(4 * S, device=dev)]
            check(lib.ssasr_blstm_bwd_bf16(ptr(x), n_rows, K, ptr(wihT_bf), ptr(whhT_p), S, n_seq, n_batch, rs_seq, rs_batch,
                                           ptr(lens_dev) if time_major else None, ptr(act), ptr(hout), ptr(cbuf),
                                           ptr(dhout), ptr(dx), ptr(dwih_p), ptr(dbias_p), ptr(dwhh_p), ptr(dcs), ptr(bar),
                                           d1 if time_major else 0, Rp, ptr(ws[0]), ptr(ws[1]), ptr(ws[2]), ptr(ws[3]),
                                           ptr(whhT_bf), ptr(xb_s), Kp_s, ptr(hb_s), st, side.cuda_stream if side else None),
                  'ssasr_blstm_bwd_bf16')
            ctx.fwd_bf = None
            ust = side.cuda_stream if side else st
            check(lib.ssasr_unpack_blstm_grads_set(ptr(dwih_p), ptr(dbias_p), ptr(dwhh_p), S, K, *[ptr(t) for t in g], ust),
                  'ssasr_unpack_blstm_grads_set')
            if side is not None:
                ev = torch.cuda.Event()
                ev.record(side)
                # NOT `g`: AccumulateGrad only adopts a gradient buffer it holds the sole reference to (otherwise it clones it --
                # on the main stream, before the side stream has written it)
                _OVERLAP['pending'].append((ev, (ws, xb_s, hb_s, dwih_p, dbias_p, dwhh_p, wihT_bf, whhT_bf)))
                _OVERLAP['deferred_total'] += 1
                _join_after_backward()
            return (dx, None, None, None) + tuple(g)
        else:
            check(lib.ssasr_blstm_bwd_f32(ptr(x), n_rows, K, ptr(wih_p), ptr(whhT_p), S, n_seq, n_batch, rs_seq, rs_batch,
                                          ptr(lens_dev) if time_major else None, ptr(act), ptr(hout), ptr(cbuf), ptr(dhout),
                                          ptr(dx), ptr(dwih_p), ptr(dbias_p), ptr(dwhh_p), ptr(dcs), ptr(bar),
                                          d1 if time_major else 0, st), 'ssasr_blstm_bwd_f32')
        g = [torch.zeros(4 * S, K, device=dev), torch.zeros(4 * S, S, device=dev), torch.zeros(4 * S, device=dev),
             torch.zeros(4 * S, device=dev), torch.zeros(4 * S, K, device=dev), torch.zeros(4 * S, S, device=dev),
             torch.zeros(4 * S, device=dev), torch.zeros(4 * S, device=dev)]
        check(lib.ssasr_unpack_blstm_grads(ptr(dwih_p), ptr(dbias_p), ptr(dwhh_p), S, K, *[ptr(t) for t in g], st),
              'ssasr_unpack_blstm_grads')
        return (dx, None, None, None) + tuple(g)


def _on_device_of(t):
    """Kernels are launched on the CURRENT stream of the tensor's device: make that device current for the call (a model on
    cuda:1 while cuda:0 is current would otherwise launch into the wrong context)."""
    _lib.require_cuda(t, 'ss_asr_b200')
    return torch.cuda.device(t.device)


def blstm(x, lens_dev, time_major, params, precision='fp32'):
    with _on_device_of(x):
        return _BLSTM.apply(x, lens_dev, time_major, precision, *params)


# --------------------------------------------------------------------------------------------------
# attend-and-spell loop
# --------------------------------------------------------------------------------------------------
class _Spell(torch.autograd.Function):
    """U steps of attention + 2 LSTM cells + character projection (asr.py:65-110)."""

    @staticmethod
    def forward(ctx, enc, enc_lens_dev, tok_in, step_mode, seed, precision, lm, phi_w, psi_w, psi_b, w_ih1, w_hh1, b_ih1, b_hh1, w_ih2,
                w_hh2, b_ih2, b_hh2, emb_w, wc, bc):
        lib = _lib.load()
        _lib.require_cuda(enc, 'Speller')
        ctx.param_refs = (phi_w, psi_w, psi_b, w_ih1, w_hh1, b_ih1, b_hh1, w_ih2, w_hh2, b_ih2, b_hh2, emb_w, wc, bc)
        skip_final = 0
        stop_token, stop_every = 0, 0
        if isinstance(lm, dict):          # {'lm': (weights, weight) or None, 'need_logits': bool, 'stop_token', 'stop_every'}
            skip_final = 0 if lm.get('need_logits', True) else 1
            stop_token, stop_every = int(lm.get('stop_token', 0)), int(lm.get('stop_every', 0))
            lm = lm.get('lm')
        enc = _f32c(enc)
        B, Tp, E = enc.shape
        Sd = w_hh1.shape[1]
        M = phi_w.shape[0]
        Cc = wc.shape[0]
        U = tok_in.shape[1]
        K1, X1, X2 = Sd + E, 2 * Sd + E, 2 * Sd
        dev = enc.device
        st = stream()
        f = lambda *s: torch.empty(*s, device=dev)
        w1cat, b1, w2cat, b2 = f(4 * Sd, X1), f(4 * Sd), f(4 * Sd, X2), f(4 * Sd)
        check(lib.ssasr_pack_lstmcell(ptr(_f32c(w_ih1)), ptr(_f32c(w_hh1)), ptr(_f32c(b_ih1)), ptr(_f32c(b_hh1)), Sd, K1,
                                      ptr(w1cat), ptr(b1), st), 'ssasr_pack_lstmcell')
        check(lib.ssasr_pack_lstmcell(ptr(_f32c(w_ih2)), ptr(_f32c(w_hh2)), ptr(_f32c(b_ih2)), ptr(_f32c(b_hh2)), Sd, Sd,
                                      ptr(w2cat), ptr(b2), st), 'ssasr_pack_lstmcell')
        phi_w, psi_w, psi_b, emb_w, wc, bc = [_f32c(t) for t in (phi_w, psi_w, psi_b, emb_w, wc, bc)]
        tok_in = tok_in.to(torch.int32).contiguous().clone()
        psi, xin1, xin2 = f(B, Tp, M), f(B, U, X1), f(B, U, X2)
        act1, act2, c1, c2, h2all = f(B, U, 4 * Sd), f(B, U, 4 * Sd), f(B, U, Sd), f(B, U, Sd), f(B, U, Sd)
        q, alpha, logits = f(B, U, M), f(B, U, Tp), f(B, U, Cc)
        modes = (C.c_int * U)(*[int(m) for m in step_mode])
        bf16 = precision == 'bf16' and X1 % 8 == 0 and X2 % 8 == 0
        w1b = w2b = wsb = encb = None
        if bf16:
            bf = lambda *s: torch.empty(*s, device=dev, dtype=torch.bfloat16)
            # step-input rows for the tcgen05 gate GEMMs: [B, X1] + one [B, X2] block per step (layer-2 chain on its own stream)
            w1b, w2b, wsb = bf(4 * Sd, X1), bf(4 * Sd, X2), bf(B * X1 + (U if _DUAL_STREAM_SPELLER else 1) * B * X2)
            encb = bf((B * Tp + M) * E)
            check(lib.ssasr_cvt_bf16(ptr(w1cat), X1, ptr(w1b), X1, 4 * Sd, X1, st), 'ssasr_cvt_bf16')
            check(lib.ssasr_cvt_bf16(ptr(w2cat), X2, ptr(w2b), X2, 4 * Sd, X2, st), 'ssasr_cvt_bf16')
        ctx.bf16 = bf16
        x3ws = None
        if precision == 'tf32x3' and not torch.is_grad_enabled() and X1 % 4 == 0 and X2 % 4 == 0:
            x3ws = torch.empty(2 * B * (X1 + X2) + 8 * Sd * (X1 + X2), device=dev)
        cl_ws = None
        if bf16 and _DUAL_STREAM_SPELLER and _CLUSTER_SPELLER:
            nb = int(lib.ssasr_speller_cl_ws_bytes(B, Tp, E, Sd, M, Cc, U))
            if nb > 0:       # cluster-persistent decoder-step kernel (spell_cl.cu): P = enc W_ctx^T, psi~, phi in bf16, G_emb
                cl_ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        lmk = {}
        if lm is not None:      # (dict of transposed fp32 tensors, weight): greedy decode with the character LM
            lmt, lm_weight = lm
            H = lmt['emb'].shape[1]
            lm_state = [torch.zeros(B, H, device=dev), torch.zeros(B, H, device=dev)]
            lmk = dict(lm_H=H, lm_weight=float(lm_weight), lm_h1=ptr(lm_state[0]), lm_h2=ptr(lm_state[1]),
                       **{'lm_' + k: ptr(v) for k, v in lmt.items()})
        steps_run = C.c_int(U)
        stop_scr = torch.zeros(1, dtype=torch.int32, device=dev) if stop_every else None     # alive until the call has returned
        a = _lib.SpellerFwdArgs(B=B, Tp=Tp, E=E, Sd=Sd, M=M, C=Cc, U=U, phi_w=ptr(phi_w), psi_w=ptr(psi_w),
                                psi_b=ptr(psi_b), w1cat=ptr(w1cat), b1=ptr(b1), w2cat=ptr(w2cat), b2=ptr(b2),
                                emb_w=ptr(emb_w), wc=ptr(wc), bc=ptr(bc), enc=ptr(enc), enc_lens=ptr(enc_lens_dev),
                                tok_in=ptr(tok_in), step_mode=C.cast(modes, C.c_void_p), seed=int(seed), psi=ptr(psi),
                                xin1=ptr(xin1), xin2=ptr(xin2), act1=ptr(act1), act2=ptr(act2), c1=ptr(c1), c2=ptr(c2),
                                h2all=ptr(h2all), q=ptr(q), alpha=ptr(alpha), logits=ptr(logits), w1cat_bf=ptr(w1b),
                                w2cat_bf=ptr(w2b), ws_bf=ptr(wsb), enc_bf=ptr(encb), x3_ws=ptr(x3ws),
                                skip_final_logits=skip_final, dual_stream=int(bf16 and _DUAL_STREAM_SPELLER),
                                stop_token=stop_token, stop_check_every=stop_every if skip_final else 0,
                                stop_scratch=ptr(stop_scr),
                                steps_run=C.addressof(steps_run), cl_ws=ptr(cl_ws), cl_ws_bytes=cl_ws.numel() if cl_ws is not None else 0,
                                **lmk)
        ctx.dual = bool(bf16 and _DUAL_STREAM_SPELLER)
        check(lib.ssasr_speller_fwd_f32(C.byref(a), st), 'ssasr_speller_fwd_f32')
        LAST_SPELL['steps_run'] = int(steps_run.value)
        ctx.save_for_backward(enc, enc_lens_dev, tok_in, phi_w, psi_w, w1cat, w2cat, wc, psi, xin1, xin2, act1, act2, c1,
                              c2, h2all, q, alpha)
        ctx.dims = (B, Tp, E, Sd, M, Cc, U)
        ctx.cl_ws = cl_ws
        ctx.w_bf = (w1b, w2b) if cl_ws is not None else None
        ctx.mark_non_differentiable(alpha, tok_in)
        return logits, alpha, tok_in

    @staticmethod
    def backward(ctx, dlogits, _dalpha, _dtok):
        lib = _lib.load()
        _consume(ctx, 'Speller')
        (enc, enc_lens_dev, tok_in, phi_w, psi_w, w1cat, w2cat, wc, psi, xin1, xin2, act1, act2, c1, c2, h2all, q,
         alpha) = ctx.saved_tensors
        B, Tp, E, Sd, M, Cc, U = ctx.dims
        K1, X1, X2 = Sd + E, 2 * Sd + E, 2 * Sd
        dev = enc.device
        st = stream()
        f = lambda *s: torch.empty(*s, device=dev)
        dlogits = _f32c(dlogits)
        d_phi_w, d_psi_w, d_psi_b = f(M, Sd), f(M, E), f(M)
        d_w1cat, d_b1, d_w2cat, d_b2 = f(4 * Sd, X1), f(4 * Sd), f(4 * Sd, X2), f(4 * Sd)
        d_emb_w, d_wc, d_bc, denc = f(Cc, Sd), f(Cc, Sd), f(Cc), f(B, Tp, E)
        scr = [f(B, U, Sd), f(B, U, X1), f(U if ctx.dual else 1, B, X2), f(B, Sd), f(B, Sd), f(B, Sd), f(B, Tp, M), f(B, U, M), f(B, U, Tp)]   # kept alive
        w1T = w2T = wsA = wsB = None
        BUp = (B * U + 7) // 8 * 8
        BTp = (B * Tp + 7) // 8 * 8
        if ctx.bf16:
            bf = lambda *s: torch.empty(*s, device=dev, dtype=torch.bfloat16)
            w1T, w2T = bf(X1, 4 * Sd), bf(X2, 4 * Sd)
            wsA = bf(max(4 * Sd * max(BUp, B), max(M, Cc, Sd) * max(BTp, BUp)))
            wsB = bf(max(X1 * BUp, E * BTp + E * M, Sd * BUp))
            check(lib.ssasr_cvt_bf16_t(ptr(w1cat), X1, ptr(w1T), 4 * Sd, 4 * Sd, X1, 0, 0, 0, 0, st), 'ssasr_cvt_bf16_t')
            check(lib.ssasr_cvt_bf16_t(ptr(w2cat), X2, ptr(w2T), 4 * Sd, 4 * Sd, X2, 0, 0, 0, 0, st), 'ssasr_cvt_bf16_t')
        # The gradients only the optimiser reads can be left to the side stream of the deferred encoder weight gradients (same
        # opt-in, same conditions: fresh .grad so that autograd adopts the returned buffers, joined by join_deferred()); they
        # then run under the latency-bound recurrent backward of encoder.blstm_4 (16 of 148 SMs busy).  Outputs are allocated
        # BEFORE the call: whatever used that memory earlier is ordered before the side stream's first write.
        fresh = all(getattr(w, 'grad', None) is None for w in ctx.param_refs)
        side = side_stream(dev) if (ctx.dual and _OVERLAP['on'] and fresh) else None
        cl_bws = None
        clk = {}
        if ctx.cl_ws is not None and _CLUSTER_SPELLER_BWD:
            nb = int(lib.ssasr_speller_cl_bwd_ws_bytes(B, Tp, E, Sd, M, U))
            if nb > 0:
                cl_bws = torch.empty(nb, dtype=torch.uint8, device=dev)
                clk = dict(cl_ws=ptr(ctx.cl_ws), w1cat_bf=ptr(ctx.w_bf[0]), w2cat_bf=ptr(ctx.w_bf[1]), cl_ws_bwd=ptr(cl_bws),
                           cl_ws_bwd_bytes=nb)
        z = lambda *s: torch.zeros(*s, device=dev)
        g1 = [z(4 * Sd, K1), z(4 * Sd, Sd), z(4 * Sd), z(4 * Sd)]
        g2 = [z(4 * Sd, Sd), z(4 * Sd, Sd), z(4 * Sd), z(4 * Sd)]
        a = _lib.SpellerBwdArgs(B=B, Tp=Tp, E=E, Sd=Sd, M=M, C=Cc, U=U, phi_w=ptr(phi_w), psi_w=ptr(psi_w),
                                w1cat=ptr(w1cat), w2cat=ptr(w2cat), wc=ptr(wc), enc=ptr(enc), enc_lens=ptr(enc_lens_dev),
                                tok_in=ptr(tok_in), psi=ptr(psi), xin1=ptr(xin1), xin2=ptr(xin2), c1=ptr(c1), c2=ptr(c2),
                                h2all=ptr(h2all), q=ptr(q), alpha=ptr(alpha), act1=ptr(act1), act2=ptr(act2),
                                dlogits=ptr(dlogits), d_phi_w=ptr(d_phi_w), d_psi_w=ptr(d_psi_w), d_psi_b=ptr(d_psi_b),
                                d_w1cat=ptr(d_w1cat), d_b1=ptr(d_b1), d_w2cat=ptr(d_w2cat), d_b2=ptr(d_b2),
                                d_emb_w=ptr(d_emb_w), d_wc=ptr(d_wc), d_bc=ptr(d_bc), denc=ptr(denc),
                                dh2all=ptr(scr[0]), dxin1=ptr(scr[1]), dxin2=ptr(scr[2]), dc1s=ptr(scr[3]),
                                dc2s=ptr(scr[4]), dh1att=ptr(scr[5]), dpsi=ptr(scr[6]), dqpre=ptr(scr[7]), de_all=ptr(scr[8]),
                                w1catT_bf=ptr(w1T), w2catT_bf=ptr(w2T), wsA=ptr(wsA), wsB=ptr(wsB), BUp=BUp, BTp=BTp,
                                dual_stream=int(ctx.dual), wgrad_stream=side.cuda_stream if side else None, **clk)
        check(lib.ssasr_speller_bwd_f32(C.byref(a), st), 'ssasr_speller_bwd_f32')
        ust = side.cuda_stream if side else st
        check(lib.ssasr_unpack_lstmcell_grads(ptr(d_w1cat), ptr(d_b1), Sd, K1, *[ptr(t) for t in g1], ust), 'unpack1')
        check(lib.ssasr_unpack_lstmcell_grads(ptr(d_w2cat), ptr(d_b2), Sd, Sd, *[ptr(t) for t in g2], ust), 'unpack2')
        if side is not None:
            ev = torch.cuda.Event()
            ev.record(side)
            # everything the side stream still reads or writes, EXCEPT the returned gradient buffers (AccumulateGrad only adopts
            # a buffer it holds the sole reference to; otherwise it clones it on the main stream, before it has been written)
            _OVERLAP['pending'].append((ev, (ctx.saved_tensors, dlogits, scr, wsA, wsB, w1T, w2T, d_w1cat, d_b1, d_w2cat, d_b2, cl_bws,
                                             ctx.cl_ws, ctx.w_bf)))
            _OVERLAP['deferred_total'] += 1
            _join_after_backward()
        return (denc, None, None, None, None, None, None, d_phi_w, d_psi_w, d_psi_b) + tuple(g1) + tuple(g2) + (d_emb_w, d_wc, d_bc)


LAST_SPELL = {'steps_run': 0}     # decoding steps the last attend-and-spell call executed (early stop of greedy decoding)


def spell(enc, enc_lens_dev, tok_in, step_mode, seed, params, precision='fp32', lm=None, need_logits=True, stop_token=0,
          stop_every=0):
    """need_logits=False (greedy decoding): the [B,U,C] logits tensor is not recomputed after the loop (left undefined)."""
    opts = lm if need_logits else {'lm': lm, 'need_logits': False, 'stop_token': stop_token, 'stop_every': stop_every}
    with _on_device_of(enc):
        return _Spell.apply(enc, enc_lens_dev, tok_in, step_mode, seed, precision, opts, *params)


def pack_charlm(rnn_lm, device):
    """CharLM (charlm.py:5-44) parameters -> the transposed fp32 tensors the LM step kernel reads."""
    sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in rnn_lm.state_dict().items()}
    t = lambda k: sd[k].t().contiguous()
    return {'emb': sd['emb.weight'].contiguous(), 'w1i': t('layer_1.weight_ih'), 'w1h': t('layer_1.weight_hh'),
            'b1i': sd['layer_1.bias_ih'].contiguous(), 'b1h': sd['layer_1.bias_hh'].contiguous(),
            'w2i': t('layer_2.weight_ih'), 'w2h': t('layer_2.weight_hh'), 'b2i': sd['layer_2.bias_ih'].contiguous(),
            'b2h': sd['layer_2.bias_hh'].contiguous(), 'wo': t('out.weight'), 'bo': sd['out.bias'].contiguous()}


# --------------------------------------------------------------------------------------------------
# loss
# --------------------------------------------------------------------------------------------------
class _ASRLoss(torch.autograd.Function):
    """trainer.py:426-434 in one kernel, gradient included."""

    @staticmethod
    def forward(ctx, logits, y):
        lib = _lib.load()
        _lib.require_cuda(logits, 'asr_loss')
        logits = _f32c(logits)
        y = y.to(torch.int64).contiguous()
        B, U, Cc = logits.shape
        L = y.shape[1]
        loss_b = torch.empty(B, device=logits.device)
        loss = torch.empty((), device=logits.device)
        dlogits = torch.empty_like(logits)
        check(lib.ssasr_ce_loss_f32(ptr(logits), ptr(y), B, U, Cc, L, ptr(loss_b), ptr(loss), ptr(dlogits), 1.0, stream()),
              'ssasr_ce_loss_f32')
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g, None


def asr_loss(logits, y):
    """loss of ASRTrainer.exec (trainer.py:426-434): logits [B,U,C], y [B,L] (label of step t = y[:, t+1])."""
    with _on_device_of(logits):
        return _ASRLoss.apply(logits, y)


# --------------------------------------------------------------------------------------------------
# single-step module API (Attention.forward / Speller.forward called one decoding step at a time,
# text_autoencoder.py:52-94).  Not the hot path of ASR.forward, but the same kernels.
# --------------------------------------------------------------------------------------------------
def _gemm(lib, M, N, K, A, lda, akm, B, ldb, bkm, Cm, ldc, bias=None, acc=0, tanh=0):
    check(lib.ssasr_gemm_f32(M, N, K, ptr(A) if not isinstance(A, int) else A, lda, akm, ptr(B) if not isinstance(B, int) else B,
                             ldb, bkm, ptr(Cm), ldc, ptr(bias), acc, tanh, stream()), 'ssasr_gemm_f32')


class _PsiMemory(torch.autograd.Function):
    """comp_listener_feature = tanh(psi(listener_feature))   (asr.py:381)."""

    @staticmethod
    def forward(ctx, enc, psi_w, psi_b):
        lib = _lib.load()
        _lib.require_cuda(enc, 'Attention')
        enc, psi_w, psi_b = _f32c(enc), _f32c(psi_w), _f32c(psi_b)
        B, Tp, E = enc.shape
        M = psi_w.shape[0]
        out = torch.empty(B, Tp, M, device=enc.device)
        _gemm(lib, B * Tp, M, E, enc, E, 1, psi_w, E, 1, out, M, psi_b, 0, 1)
        ctx.save_for_backward(enc, psi_w, out)
        return out

    @staticmethod
    def backward(ctx, d):
        lib = _lib.load()
        enc, psi_w, out = ctx.saved_tensors
        B, Tp, E = enc.shape
        M = psi_w.shape[0]
        d = _f32c(d).clone()
        check(lib.ssasr_dtanh_mul(ptr(d), ptr(out), d.numel(), stream()), 'ssasr_dtanh_mul')
        denc = torch.empty_like(enc)
        dw = torch.empty_like(psi_w)
        db = torch.empty(M, device=enc.device)
        _gemm(lib, B * Tp, E, M, d, M, 1, psi_w, E, 0, denc, E)
        _gemm(lib, M, E, B * Tp, d, M, 0, enc, E, 0, dw, E)
        check(lib.ssasr_colsum(ptr(d), ptr(db), B * Tp, M, M, 0, stream()), 'ssasr_colsum')
        return denc, dw, db


class _AttnStep(torch.autograd.Function):
    """One call of Attention.forward after the memory is cached (asr.py:383-392)."""

    @staticmethod
    def forward(ctx, h, enc, psi_t, lens_dev, phi_w):
        lib = _lib.load()
        _lib.require_cuda(h, 'Attention')
        h, enc, psi_t, phi_w = _f32c(h), _f32c(enc), _f32c(psi_t), _f32c(phi_w)
        B, Tp, E = enc.shape
        Sd, M = h.shape[1], phi_w.shape[0]
        dev = h.device
        xrow = torch.empty(B, 2 * Sd + E, device=dev)
        q = torch.empty(B, M, device=dev)
        alpha = torch.empty(B, Tp, device=dev)
        check(lib.ssasr_attn_step_fwd(B, Tp, E, Sd, M, ptr(h), ptr(phi_w), ptr(psi_t), ptr(enc), ptr(lens_dev), ptr(xrow), ptr(q),
                                      ptr(alpha), stream()), 'ssasr_attn_step_fwd')
        ctxv = xrow[:, Sd:Sd + E].contiguous()
        ctx.save_for_backward(h, enc, psi_t, lens_dev, phi_w, q, alpha)
        return alpha, ctxv

    @staticmethod
    def backward(ctx, dalpha, dctx):
        lib = _lib.load()
        h, enc, psi_t, lens_dev, phi_w, q, alpha = ctx.saved_tensors
        B, Tp, E = enc.shape
        Sd, M = h.shape[1], phi_w.shape[0]
        dev = h.device
        f = lambda *s: torch.empty(*s, device=dev)
        de, dqpre, dh, denc, dpsi = f(B, Tp), f(B, M), f(B, Sd), f(B, Tp, E), f(B, Tp, M)
        dctx = _f32c(dctx) if dctx is not None else torch.zeros(B, E, device=dev)
        dal = _f32c(dalpha) if dalpha is not None else None
        check(lib.ssasr_attn_step_bwd(B, Tp, E, Sd, M, ptr(dctx), ptr(dal), ptr(alpha), ptr(q), ptr(phi_w), ptr(psi_t), ptr(enc),
                                      ptr(lens_dev), ptr(de), ptr(dqpre), ptr(dh), ptr(denc), ptr(dpsi), stream()),
              'ssasr_attn_step_bwd')
        dphi = f(M, Sd)
        _gemm(lib, M, Sd, B, dqpre, M, 0, h, Sd, 0, dphi, Sd)
        return dh, denc, dpsi, None, dphi


class _LSTMCell(torch.autograd.Function):
    """nn.LSTMCell(x, (h, c)) -> (h', c')   (asr.py:320-324)."""

    @staticmethod
    def forward(ctx, x, h, c, w_ih, w_hh, b_ih, b_hh):
        lib = _lib.load()
        _lib.require_cuda(x, 'Speller')
        x, h, c = _f32c(x), _f32c(h), _f32c(c)
        B, Kin = x.shape
        S = w_hh.shape[1]
        dev = x.device
        st = stream()
        wcat = torch.empty(4 * S, Kin + S, device=dev)
        bcat = torch.empty(4 * S, device=dev)
        check(lib.ssasr_pack_lstmcell(ptr(_f32c(w_ih)), ptr(_f32c(w_hh)), ptr(_f32c(b_ih)), ptr(_f32c(b_hh)), S, Kin, ptr(wcat),
                                      ptr(bcat), st), 'ssasr_pack_lstmcell')
        xin = torch.cat([x, h], dim=1).contiguous()
        act = torch.empty(B, 4 * S, device=dev)
        _gemm(lib, B, 4 * S, Kin + S, xin, Kin + S, 1, wcat, Kin + S, 1, act, 4 * S, bcat)
        h2, c2 = torch.empty(B, S, device=dev), torch.empty(B, S, device=dev)
        check(lib.ssasr_lstmcell_fwd(B, S, ptr(act), ptr(c), ptr(c2), ptr(h2), st), 'ssasr_lstmcell_fwd')
        ctx.save_for_backward(xin, wcat, act, c, c2)
        ctx.dims = (B, Kin, S)
        return h2, c2

    @staticmethod
    def backward(ctx, dh2, dc2):
        lib = _lib.load()
        xin, wcat, act, c, c2 = ctx.saved_tensors
        B, Kin, S = ctx.dims
        dev = xin.device
        st = stream()
        dh2 = _f32c(dh2) if dh2 is not None else torch.zeros(B, S, device=dev)
        dc = _f32c(dc2).clone() if dc2 is not None else torch.zeros(B, S, device=dev)
        dg = act.clone()
        check(lib.ssasr_lstmcell_bwd(B, S, ptr(dg), ptr(c2), ptr(c), ptr(dh2), ptr(dc), st), 'ssasr_lstmcell_bwd')
        X = Kin + S
        dxin = torch.empty(B, X, device=dev)
        dw = torch.empty(4 * S, X, device=dev)
        db = torch.empty(4 * S, device=dev)
        _gemm(lib, B, X, 4 * S, dg, 4 * S, 1, wcat, X, 0, dxin, X)
        _gemm(lib, 4 * S, X, B, dg, 4 * S, 0, xin, X, 0, dw, X)
        check(lib.ssasr_colsum(ptr(dg), ptr(db), B, 4 * S, 4 * S, 0, st), 'ssasr_colsum')
        g = [torch.zeros(4 * S, Kin, device=dev), torch.zeros(4 * S, S, device=dev), torch.zeros(4 * S, device=dev),
             torch.zeros(4 * S, device=dev)]
        check(lib.ssasr_unpack_lstmcell_grads(ptr(dw), ptr(db), S, Kin, *[ptr(t) for t in g], st), 'ssasr_unpack_lstmcell_grads')
        return (dxin[:, :Kin].contiguous(), dxin[:, Kin:].contiguous(), dc) + tuple(g)


def psi_memory(enc, psi_w, psi_b):
    with _on_device_of(enc):
        return _PsiMemory.apply(enc, psi_w, psi_b)


def attn_step(h, enc, psi_t, lens_dev, phi_w):
    with _on_device_of(h):
        return _AttnStep.apply(h, enc, psi_t, lens_dev, phi_w)


def lstm_cell(x, h, c, cell):
    with _on_device_of(x):
        return _LSTMCell.apply(x, h, c, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)
