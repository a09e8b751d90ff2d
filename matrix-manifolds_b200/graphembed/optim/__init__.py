from .radam import RiemannianAdam
from .rsgd import RiemannianSGD

__all__ = ['RiemannianAdam', 'RiemannianSGD']
