"""Solver.step of the reference (trainer.py:131-148) fused on the device: global-norm clipping, the NaN test that cancels
the step, and torch.optim.Adadelta's update (trainer.py:401-403: Adadelta(lr, eps=1e-8)) for ALL parameter tensors in two
kernel launches and without a host synchronisation.  State layout and `state_dict()` are those of torch.optim.Adadelta, so
an optimiser checkpoint moves freely between the two."""
import ctypes as C

import torch

from . import _lib


class FusedAdadelta(torch.optim.Adadelta):
    """Drop-in for torch.optim.Adadelta (weight_decay = 0, maximize = False) with `step_clipped`.

        grad_norm = nn.utils.clip_grad_norm_(params, grad_clip)          # trainer.py:144
        if math.isnan(grad_norm): ...skip...  else: optim.step()         # trainer.py:145-148
    becomes
        optim.step_clipped(grad_clip)                                    # device-side, no sync
        optim.last_grad_norm / optim.last_applied                        # device tensors, read them only when logging
    """

    def __init__(self, params, lr=1.0, rho=0.9, eps=1e-6, weight_decay=0):
        if weight_decay != 0:
            raise NotImplementedError('FusedAdadelta: weight_decay is not supported (the reference trains without it)')
        super().__init__(params, lr=lr, rho=rho, eps=eps, weight_decay=0)
        self.last_grad_norm = None
        self.last_applied = None
        self._scratch = None
        self._out = None

    @torch.no_grad()
    def step_clipped(self, max_norm=None, write_clipped_grads=False):
        """One fused Solver.step.  max_norm None / <= 0: no clipping.  Returns the device tensor of the total gradient norm."""
        from . import functional as _F
        _F.join_deferred()                   # deferred weight gradients (functional.set_overlap_wgrad) must have landed
        lib = _lib.load()
        entries, keep = [], []
        dev = None
        hyper = None
        for group in self.param_groups:
            h = (float(group['lr']), float(group['rho']), float(group['eps']))
            if group.get('weight_decay', 0) != 0 or group.get('maximize', False):
                raise NotImplementedError('FusedAdadelta: weight_decay / maximize are not supported')
            if hyper is None:
                hyper = h
            elif h != hyper:
                raise NotImplementedError('FusedAdadelta: all parameter groups must share lr / rho / eps')
            for p in group['params']:
                if p.grad is None:
                    continue
                _lib.require_cuda(p, 'FusedAdadelta')
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise NotImplementedError('FusedAdadelta: dense fp32 parameters and gradients only')
                st = self.state[p]
                if len(st) == 0:                       # same initial state as torch.optim.Adadelta
                    st['step'] = torch.zeros((), dtype=torch.float32)
                    st['square_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['acc_delta'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if not (p.is_contiguous() and st['square_avg'].is_contiguous() and st['acc_delta'].is_contiguous()):
                    raise NotImplementedError('FusedAdadelta: contiguous parameters only')
                keep.append(g)
                entries.append((p.data_ptr(), g.data_ptr(), st['square_avg'].data_ptr(), st['acc_delta'].data_ptr(), p.numel()))
                st['step'] += 1
                dev = p.device
        if not entries:
            return None
        arr = (_lib.OptimTensor * len(entries))(*[_lib.OptimTensor(*e) for e in entries])
        need = int(lib.ssasr_adadelta_scratch_floats(arr, len(entries)))
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != dev:
            self._scratch = torch.empty(need, device=dev)
        self._out = torch.empty(2, device=dev)
        lr, rho, eps = hyper
        _lib.check(lib.ssasr_adadelta_clip_step(arr, len(entries), lr, rho, eps, float(max_norm) if max_norm else 0.0,
                                                _lib.ptr(self._scratch), _lib.ptr(self._out), 1 if write_clipped_grads else 0,
                                                _lib.stream()), 'ssasr_adadelta_clip_step')
        self.last_grad_norm = self._out[0]
        self.last_applied = self._out[1]
        return self.last_grad_norm
