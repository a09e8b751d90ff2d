"""Objective functions on (graph target, manifold squared distance) vectors.

`QuotientLoss` (the distortion loss) and `StressLoss` keep the call signature of
the reference's graphembed/objectives.py:9-45 and run as one CUDA kernel (value,
sum-reduction and derivative): csrc/gm_api.cu::product_loss_kernel.  When they
are used through `BatchedObjective` the loss is fused into the pair kernel
instead (graphembed/modules.py).

`KLDiveregenceLoss` (spelling as in the reference, objectives.py:48-76) with the
stochastic-neighbour inference model runs as two kernels (row statistics, pair
terms: csrc/gm_objectives.cu) that also produce dKL/d mdists.
"""
import abc

import torch

from . import _lib as L
from . import _ops


class ObjectiveFunction:

    @abc.abstractmethod
    def __call__(self, gdists, mdists, *, epoch, alpha):
        pass

    def loss_spec(self, *, epoch, alpha):
        """gm_loss_t description for the fused kernels, or None if this objective cannot be fused."""
        return None


class _VectorLoss(torch.autograd.Function):

    @staticmethod
    def forward(ctx, mdists, gdists, spec):
        acc, g = _ops.product_loss([mdists], [1.0], _ops.TargetSpec.vector(gdists.to(mdists.dtype)), spec)
        ctx.save_for_backward(g)
        return acc[0].to(mdists.dtype)

    @staticmethod
    def backward(ctx, upstream):
        g, = ctx.saved_tensors
        return g * upstream, None, None


class QuotientLoss(ObjectiveFunction):
    r"""sum |m / (alpha g) - 1|  (+)  sum |alpha g / (m + 1/(epoch+1)) - 1|"""

    def __init__(self, inc_l1=True, inc_l2=True):
        if not inc_l1 and not inc_l2:
            raise ValueError('At least one of the terms must be included.')
        self.inc_l1, self.inc_l2 = inc_l1, inc_l2

    def loss_spec(self, *, epoch, alpha):
        return _ops.LossSpec(L.GM_LOSS_QUOTIENT, self.inc_l1, self.inc_l2, alpha=alpha, eps=1.0 / (epoch + 1))

    def __call__(self, gdists, mdists, *, epoch, alpha):
        return _VectorLoss.apply(mdists, gdists, self.loss_spec(epoch=epoch, alpha=alpha))

    def __str__(self):
        return 'quotient_loss'


class StressLoss(ObjectiveFunction):
    r"""sum (m - g)^2"""

    def loss_spec(self, *, epoch=None, alpha=None):
        return _ops.LossSpec(L.GM_LOSS_STRESS)

    def __call__(self, gdists, mdists, *, epoch=None, alpha=None):
        return _VectorLoss.apply(mdists, gdists, self.loss_spec())

    def __str__(self):
        return 'stress_loss'


class _SneKL(torch.autograd.Function):

    @staticmethod
    def forward(ctx, mdists, gdists, alpha, inclusive):
        acc, g = _ops.sne_kl([mdists], [1.0], gdists, alpha, inclusive)
        ctx.save_for_backward(g)
        return acc[0].to(mdists.dtype)

    @staticmethod
    def backward(ctx, upstream):
        g, = ctx.saved_tensors
        return g * upstream, None, None, None


class KLDiveregenceLoss(ObjectiveFunction):
    r"""KL(p_x || p_z) (inclusive) or KL(p_z || p_x) between the stochastic-neighbour distributions induced by
    theta_x = -alpha * gdists and theta_z = -mdists over all pairs of the batch."""

    def __init__(self, inference_model, inclusive=True):
        if inference_model == 'sste':
            raise NotImplementedError('the stochastic-spanning-tree model (SSTE) is outside the B200 hot path')
        if inference_model != 'sne':
            raise ValueError(f'Inference model {inference_model} not supported.')
        self.inference_model = inference_model
        self.inclusive = inclusive

    def __call__(self, gdists, mdists, *, epoch=None, alpha):
        return _SneKL.apply(mdists, gdists, float(alpha), bool(self.inclusive))

    def __str__(self):
        return 'kl_loss'


class Sum(ObjectiveFunction):

    def __init__(self, *fns):
        self.fns = fns

    def __call__(self, *args, **kwargs):
        return sum(fn(*args, **kwargs) for fn in self.fns)

    def __str__(self):
        return '__'.join(str(f) for f in self.fns)
