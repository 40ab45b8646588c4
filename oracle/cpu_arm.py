"""CPU legs of bench.py: the reference's own implementation of each benchmarked workload, timed on the host cores.
TEST / BASELINE INFRASTRUCTURE ONLY (imported by bench.py's `cpu_baseline` / `--impl reference` legs and by tests/).

  train   the body of ASRTrainer.exec (trainer.py:415-438: prepare_x / prepare_y, zero_grad, ASR.forward with teacher forcing,
          the loss, backward, Solver.step = clip_grad_norm_(5) + NaN-skip + Adadelta) through the UNMODIFIED reference
          modules of oracle/_ref (kind "reference"); oracle/las_port.py (kind "port") only where oracle/_ref is absent
  decode  ASRTester.exec's body (trainer.py:587-591): ASR.decode per utterance, bs=1, CharLM called as the reference does
  fbank   preprocess.log_fbank per utterance under ProcessPoolExecutor(min(12, cores)) (preprocess.py:29,70); librosa is not
          installed, so the per-utterance function is oracle/fbank_oracle.py (kind "port")

Every function returns a dict {'value', 'unit', 'cores', 'kind', 'sample', ...}."""
import os
import random
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import torch

from . import ref_shim

DIMS = dict(output_dim=50, encoder_state_size=256, decoder_state_size=256, mlp_out_size=128, feature_dim=80)


def cores():
    return os.cpu_count() or 1


def _threads():
    n = cores()
    torch.set_num_threads(n)
    return n


class _RefTrain:
    """One reference model + optimiser + the ASRTrainer.exec step body."""

    def __init__(self, tf_rate=0.9, dims=DIMS):
        asr_mod, _ = ref_shim.load()
        import ASRDataset                      # the reference's own prepare_x / prepare_y (ASRDataset.py:297-340)
        self.ds = ASRDataset
        torch.manual_seed(1)
        random.seed(1)
        self.model = asr_mod.ASR(dims['output_dim'], dims['encoder_state_size'], dims['decoder_state_size'],
                                 dims['mlp_out_size'], dims['feature_dim'], tf_rate)
        self.optim = torch.optim.Adadelta(self.model.parameters(), lr=1.0, eps=1e-8)       # trainer.py:401-403
        self.loss_metric = torch.nn.CrossEntropyLoss(ignore_index=0, reduction='none')     # trainer.py:394-395
        self.kind = 'reference'

    def step(self, x64, y64):
        """x64 [1,B,T,F] float64, y64 [1,B,L] float64: what the reference DataLoader hands over (ASRDataset.py:206-226)."""
        dev = torch.device('cpu')
        x, x_lens = self.ds.prepare_x(x64, device=dev)
        y, y_lens = self.ds.prepare_y(y64, device=dev)
        ans_len = max(y_lens) - 1
        self.optim.zero_grad()
        _, prediction, _ = self.model(x, ans_len, teacher=y, state_len=x_lens)
        label = y[:, 1:ans_len + 1].contiguous()
        b, t, c = prediction.shape
        loss = self.loss_metric(prediction.view(b * t, c), label.view(-1))
        loss = torch.sum(loss.view(b, t), dim=-1) / torch.sum(y != 0, dim=-1).to(dtype=torch.float32)
        loss = torch.mean(loss)
        loss.backward()
        grad_norm = torch.nn.utils.clip_grad_norm_(self.model.parameters(), 5)               # Solver.step, trainer.py:144-148
        if not torch.isnan(grad_norm):
            self.optim.step()
        return float(loss)


class _PortTrain:
    def __init__(self, tf_rate=0.9, dims=DIMS):
        from . import las_oracle as O
        from . import las_port as P
        sd = O.make_state_dict(seed=1, **dims)
        self.port = P.Port(sd, tf_rate=tf_rate)
        self.optim = torch.optim.Adadelta(self.port.parameters(), lr=1.0, eps=1e-8)
        self.rng = random.Random(1)
        self.kind = 'port'

    def step(self, x64, y64):
        x = x64.squeeze(0).to(torch.float32)
        lens = [int(v) for v in (x.sum(-1) != 0).sum(-1)]
        y = y64.squeeze(0).to(torch.long)
        loss, _ = self.port.train_step(x, lens, y, self.optim, rng=self.rng)
        return float(loss)


def make_trainer(tf_rate=0.9, dims=DIMS):
    return _RefTrain(tf_rate, dims) if ref_shim.available() else _PortTrain(tf_rate, dims)


def loader_batch(x, y):
    """(x [B,T,F] fp32, y [B,L] int64) -> the [1,B,T,F] / [1,B,L] float64 tensors of the reference DataLoader."""
    return x.to(torch.float64).unsqueeze(0), y.to(torch.float64).unsqueeze(0)


def train(synth_batch, steps, warmup, shape, budget_s=None, tf_rate=0.9, dims=DIMS):
    """Times `steps` train steps (after `warmup`) of a B-utterance batch of the recipe `synth_batch(B, T, F, U)`.
    shape: dict(B, T, F, U).  budget_s: when given, B is reduced (32 -> 16 -> 8 -> 4) until (steps + warmup) steps are estimated
    to fit, from one calibration step at B=4."""
    n = _threads()
    tr = make_trainer(tf_rate, dims)
    B = shape['B']
    calib = None
    if budget_s is not None:
        xb, _, yb = synth_batch(4, shape['T'], shape['F'], shape['U'])
        t0 = time.perf_counter()
        tr.step(*loader_batch(xb, yb))
        calib = (time.perf_counter() - t0) / 4
        while B > 4 and (steps + warmup) * B * calib > budget_s:
            B //= 2
    x, _, y = synth_batch(B, shape['T'], shape['F'], shape['U'])
    x64, y64 = loader_batch(x, y)
    for _ in range(warmup):
        tr.step(x64, y64)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(x64, y64)
    dt = time.perf_counter() - t0
    src = ('the UNMODIFIED reference (src/asr.py + the ASRTrainer.exec body trainer.py:415-438 incl. prepare_x/y and Solver.step) '
           'from oracle/_ref' if tr.kind == 'reference' else
           'oracle/las_port.py (same torch calls as src/asr.py + trainer.py:415-438; oracle/_ref absent)')
    return {'value': B * steps / dt, 'unit': 'utt/s', 'cores': n, 'kind': tr.kind, 'batch': B, 's_per_step': dt / steps,
            'sample': 'B=%d of the C4 recipe (T=%d,F=%d,U=%d,tf %.1f), %d step(s), %d threads, torch %s; %s'
                      % (B, shape['T'], shape['F'], shape['U'], tf_rate, steps, n, torch.__version__, src)}


def decode(xs, lens, max_utts=32, budget_s=20.0, dims=DIMS):
    """Greedy ASR.decode (asr.py:112-173, bs=1, 200-character cap) over the first utterances of `xs` [N,T,F] (padded) /
    `lens`, until `max_utts` are done or `budget_s` has elapsed.  lm_weight 0 with the (randomly initialised, as in
    ASRTester, trainer.py:567-569) CharLM still evaluated per character, as the reference does (asr.py:153-159)."""
    n = _threads()
    done, chars = 0, 0
    if ref_shim.available():
        asr_mod, charlm_mod = ref_shim.load()
        import ASRDataset
        torch.manual_seed(1)
        model = asr_mod.ASR(dims['output_dim'], dims['encoder_state_size'], dims['decoder_state_size'], dims['mlp_out_size'],
                            dims['feature_dim'], 0.9)
        model.eval()
        lm = charlm_mod.CharLM(dims['output_dim'], 128)
        lm.eval()
        mapper = ASRDataset.Mapper()
        kind = 'reference'
        with torch.no_grad():
            model.decode(xs[:1, :lens[0]], [lens[0]], lm, mapper, 0.0)          # warm-up (thread pool, allocator)
            t0 = time.perf_counter()
            for i in range(min(max_utts, xs.shape[0])):
                s = model.decode(xs[i:i + 1, :lens[i]], [lens[i]], lm, mapper, 0.0)
                done += 1
                chars += len(s)
                if time.perf_counter() - t0 > budget_s:
                    break
            dt = time.perf_counter() - t0
    else:
        from . import las_oracle as O
        from . import las_port as P
        port = P.Port(O.make_state_dict(seed=1, **dims))
        kind = 'port'
        port.decode(xs[:1, :lens[0]], [lens[0]])
        t0 = time.perf_counter()
        for i in range(min(max_utts, xs.shape[0])):
            chars += len(port.decode(xs[i:i + 1, :lens[i]], [lens[i]]))
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    return {'value': done / dt, 'unit': 'utt/s', 'cores': n, 'kind': kind, 'chars_per_s': chars / dt,
            'sample': '%d utterances of the C3 set (the longest ones first), one ASR.decode call each (bs=1, 200-char cap, '
                      'lm_weight 0), %d threads, %.1f s' % (done, n, dt)}


def _fbank_one(args):
    from . import fbank_oracle as FB
    seed, n, sr, n_mels = args
    y = (0.1 * np.random.RandomState(seed).randn(n)).astype(np.float32)
    return FB.log_fbank(y, sr, n_mels).shape[0]


def fbank(n_utt=96, n_samples=160000, sr=16000, n_mels=80):
    """log_fbank per utterance in a pool of min(12, cores) worker processes, one task per utterance, as
    preprocess.iterate_by_ids does (preprocess.py:29,62-80).  The audio is generated inside the workers (the reference's
    workers read their own wav file), so nothing large crosses the process boundary."""
    w = min(12, cores())
    with ProcessPoolExecutor(max_workers=w) as ex:
        list(ex.map(_fbank_one, [(i, n_samples, sr, n_mels) for i in range(w)]))          # workers up, numpy imported
        t0 = time.perf_counter()
        frames = list(ex.map(_fbank_one, [(1000 + i, n_samples, sr, n_mels) for i in range(n_utt)]))
        dt = time.perf_counter() - t0
    return {'value': n_utt / dt, 'unit': 'utt/s', 'cores': w, 'kind': 'port',
            'sample': '%d x %.0f s utterances @%d Hz, %d mels, oracle/fbank_oracle.py (numpy restatement of log_fbank + librosa '
                      '0.6.3; librosa itself is not installed) in a pool of %d processes, %.1f s incl. synthesising the audio'
                      % (n_utt, n_samples / sr, sr, n_mels, w, dt), 'frames': int(sum(frames))}
