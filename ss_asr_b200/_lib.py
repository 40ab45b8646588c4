"""ctypes binding of libssasr.so (include/ssasr.h).  There is NO fallback: if the CUDA library is missing
or a call fails, an exception is raised."""
import ctypes as C
import os

import torch

from . import build as _build

_LIB = None


class SpellerFwdArgs(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ('B', 'Tp', 'E', 'Sd', 'M', 'C', 'U')] +
                [(n, C.c_void_p) for n in ('phi_w', 'psi_w', 'psi_b', 'w1cat', 'b1', 'w2cat', 'b2', 'emb_w', 'wc', 'bc',
                                           'enc', 'enc_lens', 'tok_in', 'step_mode')] +
                [('seed', C.c_ulonglong)] +
                [(n, C.c_void_p) for n in ('psi', 'xin1', 'xin2', 'act1', 'act2', 'c1', 'c2', 'h2all', 'q', 'alpha',
                                           'logits', 'w1cat_bf', 'w2cat_bf', 'ws_bf', 'enc_bf')] +
                [('lm_H', C.c_int), ('lm_weight', C.c_float)] +
                [(n, C.c_void_p) for n in ('lm_emb', 'lm_w1i', 'lm_w1h', 'lm_b1i', 'lm_b1h', 'lm_w2i', 'lm_w2h', 'lm_b2i',
                                           'lm_b2h', 'lm_wo', 'lm_bo', 'lm_h1', 'lm_h2', 'x3_ws')] +
                [('skip_final_logits', C.c_int), ('dual_stream', C.c_int), ('stop_token', C.c_int), ('stop_check_every', C.c_int),
                 ('stop_scratch', C.c_void_p), ('steps_run', C.c_void_p), ('cl_ws', C.c_void_p), ('cl_ws_bytes', C.c_longlong)])


class SpellerBwdArgs(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ('B', 'Tp', 'E', 'Sd', 'M', 'C', 'U')] +
                [(n, C.c_void_p) for n in ('phi_w', 'psi_w', 'w1cat', 'w2cat', 'wc', 'enc', 'enc_lens', 'tok_in',
                                           'psi', 'xin1', 'xin2', 'c1', 'c2', 'h2all', 'q', 'alpha', 'act1', 'act2',
                                           'dlogits',
                                           'd_phi_w', 'd_psi_w', 'd_psi_b', 'd_w1cat', 'd_b1', 'd_w2cat', 'd_b2',
                                           'd_emb_w', 'd_wc', 'd_bc', 'denc',
                                           'dh2all', 'dxin1', 'dxin2', 'dc1s', 'dc2s', 'dh1att', 'dpsi', 'dqpre', 'de_all',
                                           'w1catT_bf', 'w2catT_bf', 'wsA', 'wsB')] +
                [('BUp', C.c_longlong), ('BTp', C.c_longlong), ('dual_stream', C.c_int), ('wgrad_stream', C.c_void_p),
                 ('cl_ws', C.c_void_p), ('w1cat_bf', C.c_void_p), ('w2cat_bf', C.c_void_p), ('cl_ws_bwd', C.c_void_p),
                 ('cl_ws_bwd_bytes', C.c_longlong)])


class OptimTensor(C.Structure):
    _fields_ = [('p', C.c_void_p), ('g', C.c_void_p), ('sq', C.c_void_p), ('acc', C.c_void_p), ('n', C.c_longlong)]


_P, _I, _LL, _F = C.c_void_p, C.c_int, C.c_longlong, C.c_float
SIGNATURES = {
    'ssasr_last_error': (C.c_char_p, []),
    'ssasr_abi_version': (_I, []),
    'ssasr_fbank_num_frames': (_LL, [_LL, _I]),
    'ssasr_fbank': (_I, [_P, _P, _I, _I, _I, _P, _P, _I, _P]),
    'ssasr_gemm_f32': (_I, [_I, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _P, _I, _I, _P]),
    'ssasr_pack_blstm': (_I, [_P] * 8 + [_I, _I] + [_P] * 5),
    'ssasr_unpack_blstm_grads': (_I, [_P] * 3 + [_I, _I] + [_P] * 9),
    'ssasr_pack_blstm_bf16': (_I, [_P] * 8 + [_I, _I, _I] + [_P] * 6),
    'ssasr_unpack_blstm_grads_set': (_I, [_P] * 3 + [_I, _I] + [_P] * 9),
    'ssasr_pack_lstmcell': (_I, [_P] * 4 + [_I, _I] + [_P] * 3),
    'ssasr_unpack_lstmcell_grads': (_I, [_P, _P, _I, _I] + [_P] * 5),
    'ssasr_blstm_fwd_f32': (_I, [_P, _I, _I, _P, _P, _P, _I, _I, _I, _LL, _LL, _P, _P, _P, _P, _P, _P, _P, _P]),
    'ssasr_blstm_bwd_f32': (_I, [_P, _I, _I, _P, _P, _I, _I, _I, _LL, _LL, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                 _I, _P]),
    'ssasr_gemm_bf16_tc': (_I, [_I, _I, _I, _P, _LL, _I, _P, _LL, _I, _P, _I, _P, _I, _P]),
    'ssasr_gemm_tf32x3': (_I, [_I, _I, _I, _P, _P, _LL, _P, _P, _LL, _P, _I, _P, _I, _P]),
    'ssasr_gemm_bf16_tc_tn': (_I, [_I, _I, _I, _P, _LL, _I, _P, _LL, _I, _P, _I, _I, _P]),
    'ssasr_cvt_bf16': (_I, [_P, _LL, _P, _LL, _LL, _I, _P]),
    'ssasr_cvt_bf16_t': (_I, [_P, _LL, _P, _LL, _LL, _I, _I, _I, _I, _I, _P]),
    'ssasr_blstm_fwd_bf16': (_I, [_P, _I, _I, _I, _P, _P, _P, _I, _I, _I, _LL, _LL, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'ssasr_blstm_bwd_bf16': (_I, [_P, _I, _I, _P, _P, _I, _I, _I, _LL, _LL, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                  _I, _LL, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P]),
    'ssasr_speller_fwd_f32': (_I, [C.POINTER(SpellerFwdArgs), _P]),
    'ssasr_speller_bwd_f32': (_I, [C.POINTER(SpellerBwdArgs), _P]),
    'ssasr_speller_cl_ws_bytes': (_LL, [_I, _I, _I, _I, _I, _I, _I]),
    'ssasr_speller_cl_bwd_ws_bytes': (_LL, [_I, _I, _I, _I, _I, _I]),
    'ssasr_attn_step_fwd': (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'ssasr_attn_step_bwd': (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'ssasr_lstmcell_fwd': (_I, [_I, _I, _P, _P, _P, _P, _P]),
    'ssasr_lstmcell_bwd': (_I, [_I, _I, _P, _P, _P, _P, _P, _P]),
    'ssasr_dtanh_mul': (_I, [_P, _P, _LL, _P]),
    'ssasr_colsum': (_I, [_P, _P, _I, _I, _I, _I, _P]),
    'ssasr_ce_loss_f32': (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _F, _P]),
    'ssasr_adadelta_scratch_floats': (_LL, [C.POINTER(OptimTensor), _I]),
    'ssasr_adadelta_clip_step': (_I, [C.POINTER(OptimTensor), _I, _F, _F, _F, _F, _P, _P, _I, _P]),
    'ssasr_prepare_x': (_I, [_P, _I, _LL, _LL, _I, _P, _P, _P]),
    'ssasr_calc_acc_err': (_I, [_P, _LL, _LL, _I, _I, _I, _P, _LL, _I, _I, _I, _I, _P, _P, _P]),
    'ssasr_num_families': (_I, []),
    'ssasr_family_name': (C.c_char_p, [_I]),
    'ssasr_memcpy2d_h2d': (_I, [_P, _LL, _P, _LL, _LL, _LL, _P]),
    'ssasr_beam_select': (_I, [_P, _P, _F, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    'ssasr_gather_rows': (_I, [_P, _P, _P, _LL, _LL, _I, _P]),
    'ssasr_launch_count': (_LL, []),
    'ssasr_launch_count_reset': (None, []),
    'ssasr_profile_enable': (None, [_I]),
    'ssasr_profile_read': (_I, [_P, _P]),
    'ssasr_rec_tc_set_debug': (None, [_P]),
    'ssasr_rec_cl_set_debug': (None, [_P]),
    'ssasr_rec_cl_capacity': (_I, [_I, _I]),
    'ssasr_rec_cl_enable': (None, [_I]),
    'ssasr_rec_q_set_rows': (None, [_I]),
    'ssasr_rec_wide_set_debug': (None, [_P]),
    'ssasr_rec_set_dsmem': (None, [_I]),
    'ssasr_spell_cl_set_debug': (None, [_P]),
    'ssasr_spell_cl_set_debug_bwd': (None, [_P]),
    'ssasr_spell_cl_set_debug_mode': (None, [_I]),
}


def lib_path():
    return _build.LIB


def load(build_if_needed=True):
    """Loads (building first if the sources are newer) libssasr.so and types every entry point."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if build_if_needed and os.environ.get('SSASR_NO_BUILD') != '1':
        _build.build()          # a failed build of stale sources raises: an old .so would have the wrong struct layouts
    if not os.path.isfile(_build.LIB):
        raise RuntimeError('libssasr.so is missing (%s): build it with `python -m ss_asr_b200.build`; there is no '
                           'CPU/PyTorch fallback for the hot path' % _build.LIB)
    lib = C.CDLL(_build.LIB)
    try:
        lib.ssasr_abi_version.restype = C.c_int
        have = int(lib.ssasr_abi_version())
    except AttributeError:
        have = -1
    if have != _build.ABI_VERSION:
        raise RuntimeError('libssasr.so has ABI version %d, this binding needs %d: rebuild with `python -m ss_asr_b200.build '
                           '--force`' % (have, _build.ABI_VERSION))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().ssasr_last_error()
        raise RuntimeError('%s failed (rc=%d): %s' % (what, rc, msg.decode() if msg else '?'))


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), 'kernel arguments must be contiguous CUDA tensors'
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError('%s: the ss_asr_b200 hot path runs on a CUDA device only (got a %s tensor); there is '
                           'no CPU fallback' % (what, t.device))


def profile_read():
    """-> {family: (total_ms, launches)} since the last read (synchronises the device)."""
    lib = load()
    n = lib.ssasr_num_families()
    ms = (C.c_double * n)()
    cnt = (C.c_longlong * n)()
    check(lib.ssasr_profile_read(C.cast(ms, C.c_void_p), C.cast(cnt, C.c_void_p)), 'ssasr_profile_read')
    return {lib.ssasr_family_name(i).decode(): (ms[i], cnt[i]) for i in range(n)}
