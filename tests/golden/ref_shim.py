"""Import shim for the UNMODIFIED reference (cadia-lvl/ss_asr) under torch 2.x.

Only usable where /root/reference exists (the authoring container).  It is used
by make_golden.py to produce the committed fixtures and by the optional
`test_oracle_vs_reference_live` test; nothing on the GPU box imports it.

The reference cannot be imported raw: src/asr.py:12 imports a
symbol that does not exist, postprocess/preprocess pull in packages that are not
installed, and src/asr.py:378-387 builds a uint8 mask torch>=2 rejects.  The shim
stubs those and leaves /root/reference/src/*.py byte-identical.
"""
import os
import sys
import types

REF_SRC = os.environ.get("SS_ASR_REF", "/root/reference/src")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "asr.py"))


def _mod(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


def _lev(a, b):
    a, b = list(a), list(b)
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


_loaded = None


def load():
    """Returns (asr_module, charlm_module) of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference sources not present at %s" % REF_SRC)
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    for name in ("editdistance",):
        if name not in sys.modules:
            _mod(name, eval=_lev)
    for name, kw in (("librosa", {}), ("librosa.core", dict(load=None, power_to_db=None)),
                     ("librosa.feature", dict(melspectrogram=None)),
                     ("librosa.display", dict(specshow=None)),
                     ("matplotlib", {}), ("matplotlib.pyplot", {})):
        if name not in sys.modules:
            _mod(name, **kw)

    class _SW:
        def __init__(self, d):
            self.rec = []

        def add_scalar(self, k, v, s):
            self.rec.append((k, float(v), s))

        def __getattr__(self, k):
            return lambda *a, **kw: None

    if "tensorboardX" not in sys.modules:
        _mod("tensorboardX", SummaryWriter=_SW)
    import postprocess
    postprocess.Hypothesis = object
    import torch
    if not getattr(torch.Tensor.masked_fill_, "_ssasr_shim", False):
        _mf = torch.Tensor.masked_fill_

        def _masked_fill_(self, mask, value):
            return _mf(self, mask.bool() if mask.dtype == torch.uint8 else mask, value)
        _masked_fill_._ssasr_shim = True
        torch.Tensor.masked_fill_ = _masked_fill_
    import asr
    import charlm
    _loaded = (asr, charlm)
    return _loaded
