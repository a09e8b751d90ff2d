from .base import Manifold
from .euclidean import Euclidean
from .grassmann import Grassmann
from .lorentz import Lorentz
from .spd import SymmetricPositiveDefinite
from .sphere import Sphere
from .universal import Universal

__all__ = ['Manifold', 'Euclidean', 'Grassmann', 'Lorentz', 'SymmetricPositiveDefinite', 'Sphere', 'Universal']
