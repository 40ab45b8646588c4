"""Generates tests/golden/postprocess.npz by running the UNMODIFIED reference postprocess.calc_acc / calc_err with
ASRDataset.Mapper (through ref_shim; `editdistance` is stubbed by the shim's Levenshtein) on the seeded cases of
oracle/postprocess_oracle.synth_cases.

Run in the authoring container only:   python tests/golden/make_golden_postprocess.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_shim  # noqa: E402
from oracle import postprocess_oracle as PO  # noqa: E402

CASES = [dict(seed=0, B=24, U=37, L=17), dict(seed=1, B=16, U=9, L=21), dict(seed=2, B=8, U=12, L=12),
         dict(seed=3, B=5, U=1, L=1)]


def main():
    ref_shim.load(allow_container_reference=True)
    import ASRDataset
    import postprocess
    mapper = ASRDataset.Mapper()
    out = {}
    for k, c in enumerate(CASES):
        predict, label, _ = PO.synth_cases(**c)
        pt, lt = torch.from_numpy(predict), torch.from_numpy(label)
        out['acc_%d' % k] = np.float64(postprocess.calc_acc(pt, lt))
        out['err_%d' % k] = np.float64(postprocess.calc_err(pt, lt, mapper))
        # per-utterance values through the same reference functions (batch of one)
        out['acc_utt_%d' % k] = np.array([postprocess.calc_acc(pt[b:b + 1], lt[b:b + 1]) for b in range(len(lt))])
        out['err_utt_%d' % k] = np.array([postprocess.calc_err(pt[b:b + 1], lt[b:b + 1], mapper) for b in range(len(lt))])
        out['hyp_%d' % k] = np.array([mapper.translate(p) for p in np.argmax(predict, axis=-1)])
        out['case_%d' % k] = np.array([c['seed'], c['B'], c['U'], c['L']])
    np.savez_compressed(os.path.join(HERE, 'postprocess.npz'), **out)
    print({k: (v.tolist() if v.ndim == 0 else v.shape) for k, v in out.items()})


if __name__ == '__main__':
    main()
