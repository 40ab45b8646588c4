"""fbank kernel against the numpy oracle on ragged / edge-case utterances: per-utterance max error and where it is."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ss_asr_b200 import preprocess as PP
from oracle import fbank_oracle as FB
rng = np.random.RandomState(3)
ns = [16000, 8000, 12345, 401, 201, 3200, 160000, 1599, 1600, 1601, 10240, 10239, 10400, 2559, 2560]
ys = [(0.1 * rng.randn(n)).astype(np.float32) for n in ns]
ys[3][:] = 0.0
for nm in (80, 40, 23):
    outs = PP.log_fbank_batch(ys, 16000, nm)
    for n, y, o in zip(ns, ys, outs):
        w = FB.log_fbank(y, 16000, nm)
        e = np.abs(o - w)
        f, m = np.unravel_index(int(e.argmax()), e.shape)
        print(nm, n, o.shape, w.shape, 'max err %.3g at frame %d mel %d (got %.5f want %.5f)' % (e.max(), f, m, o[f, m], w[f, m]),
              'bad frames', sorted(set(np.nonzero(e > 1e-4)[0].tolist()))[:12])
