"""Recipe for `oracle/_ref/`: a byte-identical, git-ignored copy of the reference's own Python sources, so that the
UNMODIFIED reference can be executed as the CPU arm (`bench.py --impl reference`, `cpu_baseline.kind = "reference"`) and
as the drop-in witness (tests/test_dropin_trainer.py) on the GPU box, where /root/reference does not exist.

The reference is pure Python (no compiled code): "building" it is copying `src/*.py` and `conf/default.yaml` from where they
lie under /root/reference into oracle/_ref/ (listed in .gitignore -- it never enters the history -- but NOT in
.gpurunignore, so it travels with the snapshot like the built libssasr.so).  A MANIFEST with the sha256 of every file is
written next to them; `verify()` re-checks the copy against it.  TEST / BASELINE INFRASTRUCTURE ONLY: nothing under
ss_asr_b200/ reads oracle/_ref.

    python -m oracle.make_ref            # (re)create oracle/_ref from /root/reference
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
SRC_ROOT = os.environ.get('SS_ASR_REF_ROOT', '/root/reference')
FILES = [('src', n) for n in ('asr.py', 'charlm.py', 'postprocess.py', 'preprocess.py', 'ASRDataset.py', 'trainer.py',
                              'TrackerHandler.py', 'LogHandler.py', 'LMDataset.py', 'text_autoencoder.py',
                              'speech_autoencoder.py', 'discriminator.py', 'train.py')] + [('conf', 'default.yaml')]


def _sha(path):
    h = hashlib.sha256()
    with open(path, 'rb') as f:
        h.update(f.read())
    return h.hexdigest()


def make(verbose=False):
    """Copies the reference sources into oracle/_ref (no-op, returning False, where /root/reference is absent)."""
    if not os.path.isfile(os.path.join(SRC_ROOT, 'src', 'asr.py')):
        return False
    manifest = {}
    for sub, name in FILES:
        src = os.path.join(SRC_ROOT, sub, name)
        if not os.path.isfile(src):
            continue
        os.makedirs(os.path.join(DST, sub), exist_ok=True)
        dst = os.path.join(DST, sub, name)
        shutil.copyfile(src, dst)
        manifest[sub + '/' + name] = _sha(dst)
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': SRC_ROOT, 'sha256': manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print('oracle/_ref: %d reference files copied from %s' % (len(manifest), SRC_ROOT))
    return True


def verify():
    """True when oracle/_ref exists and every file still has the sha256 recorded at copy time (i.e. it is unmodified)."""
    mpath = os.path.join(DST, 'MANIFEST.json')
    if not os.path.isfile(mpath):
        return False
    man = json.load(open(mpath))['sha256']
    return all(os.path.isfile(os.path.join(DST, k)) and _sha(os.path.join(DST, k)) == v for k, v in man.items())


if __name__ == '__main__':
    ok = make(verbose=True)
    print('verify:', verify())
    sys.exit(0 if ok else 1)
