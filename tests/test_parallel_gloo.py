"""world_size-2 gloo test (CPU) of the data-parallel gradient exchange (ss_asr_b200/parallel.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Toy(torch.nn.Module):
    """Parameter names shaped like the ASR module so every bucket is exercised."""

    def __init__(self):
        super().__init__()
        self.encoder = torch.nn.ModuleDict({'blstm_%d' % i: torch.nn.Linear(4, 4) for i in range(1, 5)})
        self.embed = torch.nn.Embedding(5, 4)
        self.char_trans = torch.nn.Linear(4, 3)
        self.unused = torch.nn.Parameter(torch.zeros(2))

    def forward(self, x, idx):
        for i in range(1, 5):
            x = torch.tanh(self.encoder['blstm_%d' % i](x))
        return self.char_trans(x + self.embed(idx)).pow(2).mean()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from ss_asr_b200.parallel import GradSync, bucket_name
    torch.manual_seed(0)
    m = _Toy()
    sync = GradSync(m, world)
    assert bucket_name('encoder.blstm_3.layer.weight_ih_l0') == 'encoder.blstm_3' and bucket_name('embed.weight') == 'decoder'
    g = torch.Generator().manual_seed(100 + rank)
    for it in range(2):                                    # two steps: hooks must re-arm
        x, idx = torch.randn(6, 4, generator=g), torch.randint(0, 5, (6,), generator=g)
        for p in m.parameters():
            p.grad = None
        sync.backward(m(x, idx))
        local = _Toy()
        local.load_state_dict(m.state_dict())
        local(x, idx).backward()
        mine = torch.cat([p.grad.reshape(-1) for n, p in local.named_parameters() if p.grad is not None])
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        want = sum(gathered) / world
        used = {n for n, p in local.named_parameters() if p.grad is not None}
        got = torch.cat([p.grad.reshape(-1) for n, p in m.named_parameters() if n in used])
        assert m.unused.grad is not None and float(m.unused.grad.abs().sum()) == 0.0
        assert torch.allclose(got, want, atol=1e-7), (rank, it)
    if rank == 0:
        open(out, 'w').write('ok')
    dist.destroy_process_group()


def test_gradsync_world2_gloo(tmp_path):
    out = str(tmp_path / 'ok')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == 'ok'


def test_gradsync_world1_is_plain_backward():
    from ss_asr_b200.parallel import GradSync
    m = _Toy()
    s = GradSync(m, 1)
    s.backward(m(torch.randn(3, 4), torch.tensor([0, 1, 2])))
    assert m.char_trans.weight.grad is not None and m.unused.grad is None
