"""Riemannian Adam with a per-point scalar second moment -- the update rule of
the reference's graphembed/optim/radam.py:43-98, executed as ONE fused in-place
kernel per parameter tensor (egrad2rgrad, norm, clip, moments, exp|retr,
transport): csrc/gm_pointops.cuh::optim_update."""
import logging

import torch

from .. import _lib as L
from ._common import fused_step

logger = logging.getLogger(__name__)


class RiemannianAdam(torch.optim.Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), nc=False, max_grad_norm=None, exact=False):
        if nc and betas[1] is not None:
            logger.warning('beta2=%.5f will be ignored because `nc` is True', betas[1])
        super().__init__(params, dict(lr=lr, betas=betas, nc=nc, max_grad_norm=max_grad_norm, exact=exact))

    def _kernel_args(self, group, x):
        """(gm_optim_t, exp_avg, exp_avg_sq) for the next update of `x`; creates the state on first use."""
        beta1, beta2 = group['betas']
        clip = group['max_grad_norm']
        state = self.state[x]
        if len(state) == 0:
            state['step'] = 1  # bias correction starts at t = 1 (radam.py:56)
            state['exp_avg'] = torch.zeros_like(x, memory_format=torch.contiguous_format)
            state['exp_avg_sq'] = torch.zeros_like(x, memory_format=torch.contiguous_format)
        t = state['step']
        b2 = 1 - 1 / t if group['nc'] else beta2  # AdamNc (radam.py:82-83)
        cfg = L.Optim(kind=L.GM_OPT_RADAM, exact=int(bool(group['exact'])), has_clip=int(clip is not None),
                      step=t, has_momentum=0, first_step=int(t == 1), grassmann_retr_qr=0, zero_grad=0,
                      lr=group['lr'], beta1=beta1, beta2=b2, momentum=0.0, dampening=0.0,
                      max_grad_norm=float(clip) if clip is not None else 0.0, eps=1e-8)
        return cfg, state['exp_avg'], state['exp_avg_sq']

    def _advance(self, x, n_steps=1):
        self.state[x]['step'] += n_steps

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
                cfg, exp_avg, exp_avg_sq = self._kernel_args(group, x)
                fused_step(x, x.grad, cfg, exp_avg, exp_avg_sq)
                self._advance(x)
        return loss
