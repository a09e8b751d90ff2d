"""Riemannian SGD (optional momentum with transport) -- the update rule of the
reference's graphembed/optim/rsgd.py:40-82 as one fused in-place kernel."""
import torch
from torch.optim.optimizer import required

from .. import _lib as L
from ._common import fused_step


class RiemannianSGD(torch.optim.Optimizer):

    def __init__(self, params, lr=required, momentum=0, dampening=0, max_grad_norm=None, exact=False):
        if momentum < 0.0:
            raise ValueError('Invalid momentum value: {}'.format(momentum))
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, max_grad_norm=max_grad_norm,
                                      exact=exact))

    def _kernel_args(self, group, x):
        """(gm_optim_t, momentum buffer or None, None) for the next update of `x`; creates the state on first use."""
        clip = group['max_grad_norm']
        mom = group['momentum']
        state = self.state[x]
        first = False
        if mom > 0 and 'momentum_buffer' not in state:
            # the kernel seeds it with the Euclidean gradient on the first step (rsgd.py:53-54)
            state['momentum_buffer'] = torch.empty_like(x, memory_format=torch.contiguous_format)
            first = True
        cfg = L.Optim(kind=L.GM_OPT_RSGD, exact=int(bool(group['exact'])), has_clip=int(clip is not None),
                      step=0, has_momentum=int(mom > 0), first_step=int(first), grassmann_retr_qr=0,
                      zero_grad=0, lr=group['lr'], beta1=0.0, beta2=0.0, momentum=float(mom),
                      dampening=float(group['dampening']),
                      max_grad_norm=float(clip) if clip is not None else 0.0, eps=1e-8)
        return cfg, state.get('momentum_buffer'), None

    def _advance(self, x, n_steps=1):
        pass  # no step counter (rsgd.py:40-82)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            for x in group['params']:
                if x.grad is None:
                    continue
                cfg, buf, _ = self._kernel_args(group, x)
                fused_step(x, x.grad, cfg, buf)
        return loss
