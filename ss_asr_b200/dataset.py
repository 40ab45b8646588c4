"""Batch glue of the reference's training loop (ASRDataset.py:297-340, used at trainer.py:415-418) with the length count on
the device: `prepare_x` / `prepare_y` keep the reference signatures and return values."""
import torch

from . import _lib


def prepare_x(x, device=torch.device('cpu')):
    """ASRDataset.py:297-316.  x: [1, B, T, F] (DataLoader batch, usually float64) -> (x [B, T, F] fp32 on `device`,
    x_lens: list of the unpadded f-bank lengths = frames whose feature sum is not zero).
    On a CUDA device the cast and the count are one kernel and only the B lengths are copied back."""
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('ss_asr_b200.dataset.prepare_x: CUDA device required (no CPU fallback); got %s' % device)
    x = x.squeeze(0)
    if x.dtype not in (torch.float32, torch.float64):
        x = x.to(torch.float32)
    src = x.to(device=device, non_blocking=True).contiguous()
    B, T, F = src.shape
    lib = _lib.load()
    lens = torch.empty(B, dtype=torch.int32, device=device)
    if src.dtype == torch.float64:
        out = torch.empty(B, T, F, dtype=torch.float32, device=device)
        _lib.check(lib.ssasr_prepare_x(_lib.ptr(src), 1, B, T, F, _lib.ptr(out), _lib.ptr(lens), _lib.stream()), 'ssasr_prepare_x')
    else:
        out = src
        _lib.check(lib.ssasr_prepare_x(_lib.ptr(src), 0, B, T, F, None, _lib.ptr(lens), _lib.stream()), 'ssasr_prepare_x')
    return out, [int(v) for v in lens.tolist()]


def prepare_y(y, device=torch.device('cpu')):
    """ASRDataset.py:318-340.  y: [1, B, L] -> (y [B, L] int64 on `device`, y_lens = count(y != 0) + 1)."""
    y = y.squeeze(0).to(device=device, dtype=torch.long)
    return y, [int(v) + 1 for v in torch.sum(y != 0, dim=-1).tolist()]


class NpyFeatureStore:
    """Reader of the reference's preprocessed directory (preprocess.py:47-60,253-269): `index.tsv` with the six tab-separated
    columns `normalized_text, path_to_fbank, s_len, unpadded_num_frames, text_fname, wav_fname`, and one float64 `.npy`
    [max_len, N_DIMS] per utterance, zero-padded to the longest utterance of the whole data set.

    The reference (ASRDataset.py:91-114) loads every file whole and stacks them, i.e. a batch costs B x max_len x N_DIMS x 8
    bytes of I/O whatever its utterances' lengths.  Here a batch reads only the first `max(unpadded_num_frames)` rows of each
    file (memory-mapped), into one pinned float64 staging buffer; `batch()` then hands it to the device asynchronously and runs
    `prepare_x` there (cast + length count, ASRDataset.py:297-316), so what reaches the model is exactly what
    `prepare_x(dataset[i][0])` gives -- truncated to the batch's own longest utterance (frames beyond it are zero padding)."""

    def __init__(self, index_path):
        self.rows = []
        with open(index_path, encoding='utf-8') as f:
            for line in f:
                p = line.rstrip('\n').split('\t')
                if len(p) < 4:
                    continue
                self.rows.append({'text': p[0], 'path': p[1], 's_len': int(p[2]), 'frames': int(p[3])})
        self._stage = None

    def __len__(self):
        return len(self.rows)

    def host_batch(self, idxs):
        """-> (pinned float64 [B, Tb, F] with Tb = the longest unpadded length of the batch, [unpadded lengths], [texts])"""
        import numpy as np
        rows = [self.rows[i] for i in idxs]
        tb = max(r['frames'] for r in rows)
        maps = [np.load(r['path'], mmap_mode='r') for r in rows]
        feat = maps[0].shape[1]
        need = len(rows) * tb * feat
        if self._stage is None or self._stage.numel() < need:
            self._stage = torch.empty(need, dtype=torch.float64, pin_memory=torch.cuda.is_available())
        x = self._stage[:need].view(len(rows), tb, feat)
        xn = x.numpy()
        for i, m in enumerate(maps):
            xn[i] = m[:tb]
        return x, [r['frames'] for r in rows], [r['text'] for r in rows]

    def batch(self, idxs, device):
        """-> (x [B, Tb, F] float32 on `device`, x_lens, texts): the model input of trainer.py:415-418 for these utterances"""
        x, _, texts = self.host_batch(idxs)
        xd, lens = prepare_x(x.unsqueeze(0), device)
        return xd, lens, texts
