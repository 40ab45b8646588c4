"""CPU oracle for the log-mel filterbank frontend (numpy, fp64 where the reference is fp64).

TEST INFRASTRUCTURE ONLY (see oracle/las_oracle.py header for the import rule).

Follows /root/reference/src/preprocess.py:187-208 (`log_fbank`), whose arithmetic lives in
`librosa.feature.melspectrogram` of **librosa==0.6.3** (requirements.txt:19; FFT through
scipy==1.2.1 fftpack, requirements.txt:44).  librosa is NOT vendored under /root/reference and
not installed, and the reference holds no test vector for this function, so:

    PARITY UNPINNED by the reference itself.

The restatement below is written from librosa 0.6.3's published semantics (defaults
center=True, pad_mode='reflect', window='hann' (periodic), win_length=n_fft, power=2.0,
stft dtype=complex64, mel: fmin=0, fmax=sr/2, htk=False, norm=1 (Slaney area norm), fp64
weights, np.dot) and is cross-checked in tests/test_oracle_golden.py against two independent
implementations present in this image (torchaudio MelSpectrogram and transformers.audio_utils),
which is the best pin available offline.
"""
import numpy as np

EPS = float(np.finfo(float).eps)       # preprocess.py:201
WIN_MS, STRIDE_MS = 25, 10             # preprocess.py:31-32


def frame_params(sample_rate):
    ws = int(sample_rate * 0.001 * WIN_MS)     # preprocess.py:194
    st = int(sample_rate * 0.001 * STRIDE_MS)  # preprocess.py:195
    return ws, st


def num_frames(n_samples, sample_rate):
    ws, st = frame_params(sample_rate)
    return 1 + (n_samples + 2 * (ws // 2) - ws) // st


def hann_periodic(ws):
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(ws) / ws)


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mel = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mel)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sample_rate, n_fft, n_mels):
    """Slaney-style mel basis, [n_mels, 1 + n_fft//2], fp64 (librosa.filters.mel, 0.6.3)."""
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, float(sample_rate) / 2, n_bins, endpoint=True)
    mel_pts = np.linspace(_hz_to_mel(0.0), _hz_to_mel(sample_rate / 2.0), n_mels + 2)
    mel_f = _mel_to_hz(mel_pts)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    W = np.zeros((n_mels, n_bins))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        W[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return W * enorm[:, None]


def log_fbank(y, sample_rate, n_mels=40):
    """log_fbank(y, sample_rate) -> float32 [frames, n_mels]   (preprocess.py:187-208).
    `n_mels` is the module constant N_DIMS (preprocess.py:30, default 40; BASELINE uses 80)."""
    y = np.asarray(y)
    ws, st = frame_params(sample_rate)
    yp = np.pad(y, ws // 2, mode='reflect')
    n_frames = 1 + (len(yp) - ws) // st
    idx = np.arange(ws)[None, :] + st * np.arange(n_frames)[:, None]
    frames = yp[idx] * hann_periodic(ws)[None, :]                 # fp64 (window is fp64)
    D = np.fft.fft(frames, axis=1)[:, :1 + ws // 2].astype(np.complex64)
    P = np.abs(D) ** 2                                            # float32
    mel = mel_filterbank(sample_rate, ws, n_mels) @ P.T.astype(np.float64)   # np.dot upcasts
    out = np.log(mel + EPS).astype('float32')
    return np.swapaxes(out, 0, 1)
