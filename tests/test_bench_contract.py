"""CPU-only: the reference arm of bench.py (the one arm that runs without a GPU) prints exactly one JSON line on stdout with
the keys the driver reads; the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'asr_train_utt_per_s' and d['unit'] == 'utt/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 1 and d['steps'] == 1 and d['value'] > 0
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_product_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and 'no CUDA device' in r.stderr and r.stdout.strip() == ''


def test_decode_shards_partition_the_utterance_set():
    """decoding shards by utterance with no collective.  The per-rank shards of bench.c3_set are a partition of the
    length-sorted set (every utterance exactly once), each sorted by decreasing length (what decode_batch expects), balanced to
    within one utterance, and a rank's data do not depend on the other ranks' (same generator seeds per global index)."""
    sys.path.insert(0, ROOT)
    import bench
    n = 37
    _, full = bench.c3_set(n, 0, 1)
    assert full == sorted(full, reverse=True) and len(full) == n
    for world in (2, 4, 8):
        shards = [bench.c3_set(n, r, world) for r in range(world)]
        lens = [l for _, l in shards]
        assert sorted(sum(lens, []), reverse=True) == full
        assert max(len(l) for l in lens) - min(len(l) for l in lens) <= 1
        for xb, l in shards:
            assert l == sorted(l, reverse=True) and xb.shape == (len(l), l[0], 80)
            for i, t in enumerate(l):                       # zero padding beyond every utterance's length
                assert float(xb[i, t:].abs().sum()) == 0.0 and float(xb[i, :t].abs().sum()) > 0.0
