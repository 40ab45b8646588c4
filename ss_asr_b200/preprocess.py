"""Drop-in `log_fbank` (reference: /root/reference/src/preprocess.py:187-208) on the B200.

`log_fbank(y, sample_rate)` keeps the reference signature (numpy in, numpy [frames, N_DIMS] float32 out);
`log_fbank_batch` is the batched entry the preprocessing loop (preprocess.py:62-80) should call instead of
one process-pool task per utterance.  N_DIMS / WIN_SIZE / STRIDE are the reference's module constants
(preprocess.py:30-32)."""
import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream

N_DIMS = 40      # preprocess.py:30 (BASELINE configs override to 80)
WIN_SIZE = 25    # ms, preprocess.py:31
STRIDE = 10      # ms, preprocess.py:32
MAX_UTT_PER_CALL = 65535


def num_frames(n_samples, sample_rate):
    return int(_lib.load().ssasr_fbank_num_frames(int(n_samples), int(sample_rate)))


def log_fbank_device(audio, offsets, sample_rate, n_mels=None, out=None):
    """audio: 1-D float32 CUDA tensor holding all utterances back to back; offsets: int64 CPU tensor/list
    [n_utt+1].  Returns (fbank [total_frames, n_mels] CUDA float32, frame_offsets list)."""
    lib = _lib.load()
    _lib.require_cuda(audio, 'log_fbank')
    n_mels = N_DIMS if n_mels is None else int(n_mels)
    off = [int(v) for v in (offsets.tolist() if torch.is_tensor(offsets) else offsets)]
    n_utt = len(off) - 1
    ws = int(sample_rate * 0.001 * WIN_SIZE)
    lens = [off[i + 1] - off[i] for i in range(n_utt)]
    if min(lens) <= ws // 2:
        raise ValueError('log_fbank: every utterance needs more than %d samples (reflect padding), got %d'
                         % (ws // 2, min(lens)))
    frames = [num_frames(n, sample_rate) for n in lens]
    foff = [0]
    for f in frames:
        foff.append(foff[-1] + f)
    dev = audio.device
    if out is None:
        out = torch.empty(foff[-1], n_mels, device=dev, dtype=torch.float32)
    off_d = torch.tensor(off, dtype=torch.int64, device=dev)
    foff_d = torch.tensor(foff, dtype=torch.int64, device=dev)
    audio = audio.contiguous()
    for u0 in range(0, n_utt, MAX_UTT_PER_CALL):
        u1 = min(n_utt, u0 + MAX_UTT_PER_CALL)
        check(lib.ssasr_fbank(ptr(audio), off_d.data_ptr() + 8 * u0, u1 - u0, int(sample_rate), n_mels, ptr(out),
                              foff_d.data_ptr() + 8 * u0, max(frames[u0:u1]), stream()), 'ssasr_fbank')
    return out, foff


def log_fbank_batch(ys, sample_rate, n_mels=None, device='cuda'):
    """ys: list of 1-D numpy arrays.  Returns a list of numpy [frames, n_mels] float32 arrays."""
    off = [0]
    for y in ys:
        off.append(off[-1] + len(y))
    host = torch.from_numpy(np.concatenate([np.asarray(y, dtype=np.float32) for y in ys]))
    audio = host.pin_memory().to(device, non_blocking=True) if torch.cuda.is_available() else host
    out, foff = log_fbank_device(audio, off, sample_rate, n_mels)
    res = out.cpu().numpy()
    return [res[foff[i]:foff[i + 1]] for i in range(len(ys))]


def log_fbank(y, sample_rate):
    """Given a signal and a sample rate, the [num_frames, N_DIMS] float32 log mel filterbank (preprocess.py:187)."""
    return log_fbank_batch([y], sample_rate, N_DIMS)[0]
