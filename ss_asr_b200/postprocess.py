"""Validation metrics of the reference (src/postprocess.py:7-75) and the body of ASRTrainer.valid (src/trainer.py:472-494)
with the per-utterance work on the device: `calc_acc`, `calc_err`, `trim_eos`, `draw_att` keep the
reference signatures and return values.

The reference copies the whole [B, U, C] prediction to the host, takes the argmax with numpy and runs Python double loops
plus `editdistance.eval` per utterance.  Here one kernel (`ssasr_calc_acc_err`, csrc/postprocess.cu) produces, per utterance,
{correct, total, word edit distance, label words}; only those 4 int32 travel back and the final float arithmetic is the
reference's own (Python floats, same order), so the returned numbers are identical, not merely close."""
import torch

from . import _lib

SOS_TKN, EOS_TKN = '<', '>'            # preprocess.py:24-25


def _mapper_ids(mapper):
    """(sos_id, eos_id, space_id) of a Mapper (ASRDataset.py:228-262); the default table when mapper is None."""
    if mapper is None:
        return 0, 1, 46

    def ind(ch):
        try:
            return int(mapper.char_to_ind(ch))
        except KeyError:
            return -1
    return ind(SOS_TKN), ind(EOS_TKN), ind(' ')


def utterance_stats(predict, label, mapper=None, return_tokens=False):
    """predict [B, U, C] float (CUDA), label [B, L] integer -> int32 [B, 4] on the host:
    (correct, total) of calc_acc and (word edit distance, label words) of calc_err; optionally the argmax tokens [B, U]."""
    _lib.require_cuda(predict, 'postprocess')
    lib = _lib.load()
    predict = predict.detach()
    if predict.dtype != torch.float32:
        predict = predict.float()
    if predict.stride(-1) != 1:
        predict = predict.contiguous()
    label = label.to(device=predict.device, dtype=torch.int64)
    if label.dim() != 2 or predict.dim() != 3 or label.shape[0] != predict.shape[0]:
        raise ValueError('postprocess: predict [B,U,C] and label [B,L] expected, got %s and %s'
                         % (tuple(predict.shape), tuple(label.shape)))
    if label.stride(-1) != 1:
        label = label.contiguous()
    B, U, Cc = predict.shape
    L = label.shape[1]
    stats = torch.empty(B, 4, dtype=torch.int32, device=predict.device)
    toks = torch.empty(B, U, dtype=torch.int32, device=predict.device) if return_tokens else None
    if B and (U == 0 or L == 0):
        raise ValueError('postprocess: empty prediction or label sequence')
    sos, eos, space = _mapper_ids(mapper)
    _lib.check(lib.ssasr_calc_acc_err(predict.data_ptr(), predict.stride(0), predict.stride(1), B, U, Cc, label.data_ptr(),
                                      label.stride(0), L, sos, eos, space, _lib.ptr(stats), _lib.ptr(toks), _lib.stream()),
               'ssasr_calc_acc_err')
    stats = stats.cpu()
    return (stats, toks) if return_tokens else stats


def _acc_from(stats):
    accs = [float(c) / t for c, t, _, _ in stats.tolist()]      # ZeroDivisionError for an empty label, as the reference
    return sum(accs) / len(accs)


def _err_from(stats):
    ds = [float(d) / n for _, _, d, n in stats.tolist()]
    return sum(ds) / len(ds)


def calc_acc(predict, label):
    """postprocess.py:7-29: character accuracy up to the first 0 of each label, averaged over the batch."""
    return _acc_from(utterance_stats(predict, label))


def calc_err(predict, label, mapper):
    """postprocess.py:31-50: word-level edit distance / label words per utterance, averaged over the batch."""
    return _err_from(utterance_stats(predict, label, mapper))


def calc_acc_err(predict, label, mapper):
    """Both metrics from one launch (ASRTrainer.valid calls them back to back on the same tensors, trainer.py:493-494)."""
    stats = utterance_stats(predict, label, mapper)
    return _acc_from(stats), _err_from(stats)


def trim_eos(sequence, eos_id=1):
    """Host helper with the contract of postprocess.py:68-75: the token ids of `sequence` as Python ints up to AND INCLUDING
    the first EOS (id 1 with the default Mapper); the whole sequence when there is none."""
    ids = [int(v) for v in sequence]
    return ids[:ids.index(eos_id) + 1] if eos_id in ids else ids


def draw_att(att_maps, hyps):
    """Host helper with the contract of postprocess.py:52-66 (TensorBoard images for ASRTrainer.valid, trainer.py:512):
    per utterance a [3, n, T'] tensor -- its attention map repeated over three colour channels, cut to the n decoding steps
    up to and including the first EOS of its hypothesis."""
    out = []
    for att, hyp in zip(att_maps, hyps):
        n = len(trim_eos(hyp))
        out.append(att[:n].unsqueeze(0).expand(3, -1, -1).contiguous())
    return out


@torch.no_grad()
def valid_step(asr_model, x, y, x_lens, y_lens, mapper):
    """The per-batch body of ASRTrainer.valid (trainer.py:472-494): batched greedy forward without a teacher for
    ans_len + 30 steps, the training loss on the first ans_len steps, calc_acc and calc_err.
    Returns (loss tensor on the device, acc, err, prediction, att_map, argmax tokens [B, ans_len + 30] int32)."""
    from .functional import asr_loss
    ans_len = max(y_lens) - 1
    _, prediction, att_map = asr_model(x, ans_len + 30, state_len=x_lens)
    label = y[:, 1:ans_len + 1]
    loss = asr_loss(prediction[:, :ans_len, :], y)
    stats, toks = utterance_stats(prediction, label, mapper, return_tokens=True)
    return loss, _acc_from(stats), _err_from(stats), prediction, att_map, toks
