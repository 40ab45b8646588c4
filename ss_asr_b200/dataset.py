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
