"""Batch-sharded data parallelism for the LAS train step: one process per GPU, replicated
parameters, ONE exchange per step -- the gradient all-reduce -- issued per bucket as soon as that bucket's
backward kernel has finished, so NCCL (NVLink 5 / NVSwitch) overlaps the rest of the backward pass.

Buckets follow backward-completion order: decoder side (attention, speller, embed, char_trans) first, then
encoder.blstm_4 ... encoder.blstm_1.  Each rank runs the reference semantics on its OWN batch (blstm_4 couples
the utterances of a batch), losses are per-rank batch means, so gradients are averaged.
Works with any backend (`nccl` on the GPUs, `gloo` in the CPU tests)."""
import torch
import torch.distributed as dist

BUCKET_PREFIXES = ('encoder.blstm_1', 'encoder.blstm_2', 'encoder.blstm_3', 'encoder.blstm_4')


def bucket_name(param_name):
    for p in BUCKET_PREFIXES:
        if param_name.startswith(p):
            return p
    return 'decoder'


class GradSync:
    def __init__(self, model, world=None, group=None):
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.group = group
        self.buckets = {}
        self.pending = []
        self._handles = []
        self.enabled = True          # False: backward() leaves every rank with its own gradients (measurement of the exchange cost)
        if self.world > 1:
            for name, p in model.named_parameters():
                if not p.requires_grad:
                    continue
                b = self.buckets.setdefault(bucket_name(name), {'params': [], 'ready': 0})
                b['params'].append(p)
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(b)))

    def _make_hook(self, bucket):
        def hook(_p):
            if not self.enabled:
                return
            bucket['ready'] += 1
            if bucket['ready'] == len(bucket['params']):
                self._launch(bucket)
        return hook

    def _launch(self, bucket):
        from . import functional as Fk
        if Fk.overlap_wgrad_enabled() and bucket['params'][0].is_cuda:
            # the bucket's gradients may still be in flight on the deferred weight-gradient stream: gather and reduce THERE
            # (ordered after them) instead of making the main stream -- and with it the rest of the backward pass -- wait
            side = Fk.side_stream(bucket['params'][0].device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                flat = torch.cat([p.grad.reshape(-1) for p in bucket['params']])
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            flat = torch.cat([p.grad.reshape(-1) for p in bucket['params']])
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending.append((bucket, flat, work))

    def backward(self, loss):
        """loss.backward() with the bucketed, overlapped gradient average."""
        for b in self.buckets.values():
            b['ready'] = 0
        self.pending = []
        loss.backward()
        if self.world <= 1 or not self.enabled:
            return
        from . import functional as Fk
        Fk.join_deferred()
        for b in self.buckets.values():          # parameters that received no gradient this step
            if 0 < b['ready'] < len(b['params']) or (b['ready'] == 0 and any(p.grad is not None for p in b['params'])):
                for p in b['params']:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                if not any(pb is b for pb, _, _ in self.pending):
                    self._launch(b)
        inv = 1.0 / self.world
        for bucket, flat, work in self.pending:
            work.wait()
            if flat.is_cuda:
                flat.record_stream(torch.cuda.current_stream())
            flat.mul_(inv)                      # one launch per bucket; the gradients become views of the reduced buffer
            off = 0
            for p in bucket['params']:
                n = p.numel()
                p.grad = flat[off:off + n].view_as(p)
                off += n
        self.pending = []

    def close(self):
        for h in self._handles:
            h.remove()
        self._handles = []


class HostBatchPipeline:
    """Double-buffered host -> device staging of training batches (the DataLoader hand-off of trainer.py:415-418):
    the pinned-memory copy of batch i+1 runs on a copy stream while batch i is computing, so the H2D transfer (42 MB per
    256-utterance batch) leaves the critical path.  `submit` enqueues the copies, `take` makes the compute stream wait for
    them and returns the device tensors."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._queue = []

    def submit(self, *host_tensors):
        with torch.cuda.stream(self.stream):
            devs = [t.to(self.device, non_blocking=True) for t in host_tensors]
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._queue.append((devs, ev))

    def take(self):
        devs, ev = self._queue.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for d in devs:
            d.record_stream(cur)
        return devs
