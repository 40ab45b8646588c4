"""The reference's on-disk feature format (preprocess.py:47-60,253-269: index.tsv + zero-padded float64 .npy) read by
ss_asr_b200.dataset.NpyFeatureStore: host part against the reference's own ASRDataset, device part against its prepare_x."""
import numpy as np
import pytest
import torch

import dropin_harness as H
from oracle import ref_shim


def _ref_dataset(index, batch_size):
    """the unmodified ASRDataset (ASRDataset.py) over the same index"""
    ref_shim.load(True)
    with ref_shim.stubs():
        import importlib
        ds = importlib.import_module('ASRDataset')
    return ds, ds.ASRDataset(index, batch_size=batch_size) if hasattr(ds, 'ASRDataset') else None


def test_host_batch_reads_only_the_batch_longest_rows(tmp_path):
    from ss_asr_b200.dataset import NpyFeatureStore
    index = H.make_dataset(str(tmp_path / 'data'), n_utt=8, feat=40, t_min=20, t_max=90)
    store = NpyFeatureStore(index)
    assert len(store) == 8
    x, lens, texts = store.host_batch([4, 5, 6, 7])                 # the four shortest utterances
    full = np.stack([np.load(store.rows[i]['path']) for i in (4, 5, 6, 7)])
    assert x.dtype == torch.float64 and x.shape[1] == max(lens) < full.shape[1]
    assert np.array_equal(x.numpy(), full[:, :max(lens)])
    assert not full[:, max(lens):].any()                            # what was skipped is zero padding
    assert lens == [int((full[i].sum(-1) != 0).sum()) for i in range(4)]
    assert all(t.startswith('<') and t.endswith('>') for t in texts)
    if ref_shim.available(allow_container_reference=True):
        ds, d = _ref_dataset(index, 4)
        if d is not None and hasattr(d, 'get_batched_fbanks_by_paths'):
            ref = d.get_batched_fbanks_by_paths([store.rows[i]['path'] for i in (4, 5, 6, 7)])
            assert np.array_equal(ref[:, :max(lens)], x.numpy())


@pytest.mark.gpu
def test_device_batch_matches_reference_prepare_x(tmp_path):
    from ss_asr_b200.dataset import NpyFeatureStore
    index = H.make_dataset(str(tmp_path / 'data'), n_utt=8, feat=40, t_min=20, t_max=90)
    store = NpyFeatureStore(index)
    xd, lens, _ = store.batch([0, 1, 2, 3], 'cuda')
    full = np.stack([np.load(store.rows[i]['path']) for i in (0, 1, 2, 3)])
    # ASRDataset.py:297-316 on the whole padded files: cast to float32, lengths = frames whose feature sum is not zero
    want = torch.from_numpy(full).to(torch.float32)
    want_lens = [int(v) for v in (want.sum(-1) != 0).sum(-1)]
    assert lens == want_lens and xd.shape[1] == max(lens)
    assert torch.equal(xd.cpu(), want[:, :max(lens)])
