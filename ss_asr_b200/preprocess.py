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


class FbankPlan:
    """Device-side offsets for one batch geometry, so that repeated extraction (preprocessing loop, benchmark) costs one
    kernel launch per <= 65535 utterances and no per-call host work."""

    def __init__(self, offsets, sample_rate, n_mels=None, device='cuda'):
        self.sample_rate = int(sample_rate)
        self.n_mels = N_DIMS if n_mels is None else int(n_mels)
        off = [int(v) for v in (offsets.tolist() if torch.is_tensor(offsets) else offsets)]
        self.n_utt = len(off) - 1
        ws = int(sample_rate * 0.001 * WIN_SIZE)
        st = int(sample_rate * 0.001 * STRIDE)
        lens = [off[i + 1] - off[i] for i in range(self.n_utt)]
        if min(lens) <= ws // 2:
            raise ValueError('log_fbank: every utterance needs more than %d samples (reflect padding), got %d'
                             % (ws // 2, min(lens)))
        self.frames = [1 + (n + 2 * (ws // 2) - ws) // st for n in lens]       # == ssasr_fbank_num_frames
        self.foff = [0]
        for f in self.frames:
            self.foff.append(self.foff[-1] + f)
        self.off_d = torch.tensor(off, dtype=torch.int64, device=device)
        self.foff_d = torch.tensor(self.foff, dtype=torch.int64, device=device)
        self.n_samples = off[-1]
        self.chunks = [(u0, min(self.n_utt, u0 + MAX_UTT_PER_CALL)) for u0 in range(0, self.n_utt, MAX_UTT_PER_CALL)]
        self.chunk_max = [max(self.frames[a:b]) for a, b in self.chunks]

    def run(self, audio, out=None):
        lib = _lib.load()
        _lib.require_cuda(audio, 'log_fbank')
        if out is None:
            out = torch.empty(self.foff[-1], self.n_mels, device=audio.device, dtype=torch.float32)
        audio = audio.contiguous()
        st = stream()
        for (u0, u1), mx in zip(self.chunks, self.chunk_max):
            check(lib.ssasr_fbank(ptr(audio), self.off_d.data_ptr() + 8 * u0, u1 - u0, self.sample_rate, self.n_mels, ptr(out),
                                  self.foff_d.data_ptr() + 8 * u0, mx, st), 'ssasr_fbank')
        return out


    def run_host(self, audio_host, out_host, n_parts=8):
        """Host buffers in, host buffers out (both pinned): the utterances are processed in `n_parts` contiguous groups whose
        host -> device copy, kernel and device -> host copy run on three streams, so the PCIe transfers in both directions
        overlap each other and the kernel (what the preprocessing loop of preprocess.py:62-80 needs: waveforms come from and
        features go back to host memory).  Returns after enqueuing; `out_host` is complete after `self.sync()`."""
        lib = _lib.load()
        dev = self.off_d.device
        assert audio_host.is_pinned() and out_host.is_pinned() and audio_host.numel() == self.n_samples
        if getattr(self, '_pipe', None) is None:
            self._pipe = {'s_in': torch.cuda.Stream(dev), 's_out': torch.cuda.Stream(dev),
                          'audio': torch.empty(self.n_samples, device=dev, dtype=torch.float32),
                          'fb': torch.empty(self.foff[-1], self.n_mels, device=dev, dtype=torch.float32),
                          'computed': None, 'copied_out': None}
        pp = self._pipe
        cur = torch.cuda.current_stream(dev)
        out_flat = out_host.view(-1, self.n_mels)
        off = self.off_d_host if hasattr(self, 'off_d_host') else None
        if off is None:
            off = self.off_d_host = [int(v) for v in self.off_d.tolist()]
        if pp['computed'] is not None:
            pp['s_in'].wait_event(pp['computed'])          # the previous call's kernels have read the device audio buffer
        if pp['copied_out'] is not None:
            cur.wait_event(pp['copied_out'])               # ... and its features have left the device output buffer
        per = (self.n_utt + n_parts - 1) // n_parts
        for u0 in range(0, self.n_utt, per):
            u1 = min(self.n_utt, u0 + per)
            with torch.cuda.stream(pp['s_in']):
                pp['audio'][off[u0]:off[u1]].copy_(audio_host[off[u0]:off[u1]], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(pp['s_in'])
            cur.wait_event(ev_in)
            for c0 in range(u0, u1, MAX_UTT_PER_CALL):
                c1 = min(u1, c0 + MAX_UTT_PER_CALL)
                check(lib.ssasr_fbank(ptr(pp['audio']), self.off_d.data_ptr() + 8 * c0, c1 - c0, self.sample_rate, self.n_mels,
                                      ptr(pp['fb']), self.foff_d.data_ptr() + 8 * c0, max(self.frames[c0:c1]), cur.cuda_stream),
                      'ssasr_fbank')
            ev_k = torch.cuda.Event()
            ev_k.record(cur)
            pp['computed'] = ev_k
            with torch.cuda.stream(pp['s_out']):
                pp['s_out'].wait_event(ev_k)
                out_flat[self.foff[u0]:self.foff[u1]].copy_(pp['fb'][self.foff[u0]:self.foff[u1]], non_blocking=True)
                ev_o = torch.cuda.Event()
                ev_o.record(pp['s_out'])
                pp['copied_out'] = ev_o
        return out_host

    def sync(self):
        if getattr(self, '_pipe', None) is not None and self._pipe['copied_out'] is not None:
            self._pipe['copied_out'].synchronize()


def log_fbank_device(audio, offsets, sample_rate, n_mels=None, out=None):
    """audio: 1-D float32 CUDA tensor holding all utterances back to back; offsets: int64 CPU tensor/list
    [n_utt+1].  Returns (fbank [total_frames, n_mels] CUDA float32, frame_offsets list)."""
    _lib.require_cuda(audio, 'log_fbank')
    plan = FbankPlan(offsets, sample_rate, n_mels, device=audio.device)
    return plan.run(audio, out), plan.foff


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


def fbank_to_listener_batch(fb, frame_offsets):
    """fbank -> Listener hand-off without the `.npy` round trip (preprocess.py:47-60 writes every utterance to disk zero-padded
    to the data set's longest; ASRDataset reads it back): `fb` [total_frames, n_mels] as returned by `log_fbank_device` /
    `FbankPlan.run` stays on the device and is re-laid-out as the zero-padded batch `x [B, T_max, n_mels]` the Listener takes,
    utterances in decreasing length (the pack_padded_sequence contract, asr.py:413).
    Returns (x, lens, order) with order[i] = index of the i-th batch row in the original utterance order."""
    n = len(frame_offsets) - 1
    lens_all = [int(frame_offsets[i + 1]) - int(frame_offsets[i]) for i in range(n)]
    order = sorted(range(n), key=lambda i: -lens_all[i])
    lens = [lens_all[i] for i in order]
    x = fb.new_zeros(n, lens[0], fb.shape[1])
    for row, i in enumerate(order):                     # B device-to-device copies (data movement only)
        x[row, :lens[row]].copy_(fb[int(frame_offsets[i]):int(frame_offsets[i + 1])])
    return x, lens, order
