"""ss_asr_b200 -- B200-native (sm_100a) drop-in for the Listen-Attend-Spell hot path of cadia-lvl/ss_asr.

    from ss_asr_b200.asr import ASR, Listener, Speller, Attention, pBLSTM      # replaces src/asr.py
    from ss_asr_b200.preprocess import log_fbank                                # replaces preprocess.log_fbank
    from ss_asr_b200.functional import asr_loss                                 # fused trainer.py:426-434

Host code is Python/PyTorch (device memory, streams, autograd graph, torch.distributed); all compute is in
libssasr.so (hand-written CUDA, C ABI in include/ssasr.h).  There is no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401

__all__ = ['asr', 'preprocess', 'functional', 'build']
