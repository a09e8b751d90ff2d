"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement (torch CPU tensors + autograd, exactly the ATen/LAPACK calls the
reference makes) of the graphembed training hot path of dalab/matrix-manifolds:
manifold squared distances, the distortion losses, and the Riemannian optimizer
updates.  Every function cites the reference lines it restates
(paths relative to /root/reference/graphembed/graphembed/).

Parity status: PINNED.  tests/test_oracle_pinned.py compares every function here
with golden vectors in tests/golden/*.npz that were produced by importing and
running the real reference in the build container
(tests/golden/make_golden.py, using oracle/ref_import.py); when /root/reference is
present the same test also compares against the live reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg may import this module (as the checker / CPU baseline).
"""
import math

import torch

EPS = 1e-8  # utils.py:13 -- same constant for fp32 and fp64


# ----------------------------------------------------------------------------
# small linalg helpers (linalg/torch_batch.py)
# ----------------------------------------------------------------------------
def _sym(x):  # torch_batch.py:25-27
    return 0.5 * (x + x.transpose(-2, -1))


def _axat(a, x):  # torch_batch.py:30-34
    return torch.einsum('...ij,...jk,...lk->...il', a, x, a)


def _st_clamp_(t, lo=None, hi=None):
    """`t.data.clamp_()`: value-only clamp that autograd does not see."""
    t.data.clamp_(min=lo, max=hi)
    return t


def _eigh(x):  # torch_batch.py:127-135 (torch.symeig, upper=True)
    return torch.linalg.eigh(x, UPLO='U')


def _symapply(x, f, wmin=None, wmax=None):  # torch_batch.py:145-153
    w, v = _eigh(x)
    if wmin is not None or wmax is not None:
        _st_clamp_(w, wmin, wmax)
    return torch.einsum('...ij,...j,...kj->...ik', v, f(w), v)


# ----------------------------------------------------------------------------
# closed forms with the reference's eps terms (linalg/fast.py)
# ----------------------------------------------------------------------------
def eigvals_2x2_closed(m, eps=1e-8):  # fast.py:53-70
    a, b, c = m[..., 0, 0], m[..., 1, 1], m[..., 0, 1]
    det = a * b - c**2
    half_tr = 0.5 * (a + b)
    delta = half_tr**2 - det
    _st_clamp_(delta, eps)
    root = delta.sqrt()
    return torch.stack([half_tr - root, half_tr + root], dim=-1)


def eigvals_3x3_closed(m, eps=1e-8):  # fast.py:75-91
    q = m.diagonal(dim1=-2, dim2=-1).sum(-1).view(-1, 1, 1) / 3
    y = m - q * torch.eye(3, dtype=m.dtype).expand_as(m)
    p = torch.sqrt(y.pow(2).sum((-2, -1), keepdim=True) / 6)
    _st_clamp_(p, eps)
    y00, y01, y02 = y[..., 0, 0], y[..., 0, 1], y[..., 0, 2]
    y11, y12, y22 = y[..., 1, 1], y[..., 1, 2], y[..., 2, 2]
    det = (y00 * y11 * y22 + 2 * y01 * y02 * y12 - y11 * y02**2 - y00 * y12**2 - y22 * y01**2).view(-1, 1, 1)
    r = det / (2 * p.pow(3) + eps)
    _st_clamp_(r, -1 + eps, 1 - eps)
    phi = torch.acos(r) / 3
    e1 = q + 2 * p * torch.cos(phi)
    e2 = q + 2 * p * torch.cos(phi + 2 * math.pi / 3)
    e3 = 3 * q - e1 - e2
    return torch.stack([e2, e3, e1], dim=-1).squeeze()


def chol_2x2_closed(x, eps=1e-8, want_inverse=False):  # fast.py:94-134
    shp = x.shape[:-2] + (1, 1)
    x00 = x[..., 0, 0].reshape(shp)
    x00 = torch.where(x00.detach() < eps, x00 + (eps - x00.detach()), x00)  # value-only clamp, no aliasing
    x11 = x[..., 1, 1].reshape(shp)
    x01 = x[..., 0, 1].reshape(shp)
    a = x00.sqrt()
    b = x01 / a
    c = (x11 - b**2 + eps).sqrt()
    zero = torch.zeros_like(a)
    l = torch.cat([torch.cat([a, zero], -1), torch.cat([b, c], -1)], -2)
    if not want_inverse:
        return l
    det = a * c
    _st_clamp_(det, eps)
    l_inv = torch.cat([torch.cat([c, zero], -1), torch.cat([-b, a], -1)], -2) / det
    return l_inv, l


# ----------------------------------------------------------------------------
# SPD (manifolds/spd.py)
# ----------------------------------------------------------------------------
class SpdOracle:
    def __init__(self, n, fast_symeig=True, fast_chol=True, stein=False, wmin=1e-8, wmax=1e8):
        self.n, self.stein, self.wmin, self.wmax = n, stein, wmin, wmax
        self.fast_eig = fast_symeig and n in (2, 3)  # spd.py:35-41
        self.fast_chol = fast_chol and n == 2  # spd.py:43-49

    # spd.py:43-61
    def chol(self, x):
        return chol_2x2_closed(x) if self.fast_chol else torch.linalg.cholesky(x)

    def invchol(self, x):
        if self.fast_chol:
            return chol_2x2_closed(x, want_inverse=True)
        l = torch.linalg.cholesky(x)
        eye = torch.eye(self.n, dtype=x.dtype)
        return torch.linalg.solve_triangular(l, eye.expand_as(l), upper=False), l

    def _eigvals(self, m):
        if self.fast_eig:
            return eigvals_2x2_closed(m) if self.n == 2 else eigvals_3x3_closed(m)
        return _eigh(m)[0]  # spd.py:63-64

    def _norm_log2(self, m):  # spd.py:163-169 (squared=True)
        w = self._eigvals(m)
        w = w.reshape(m.shape[0], -1) if w.ndim == 1 and m.shape[0] == 1 else w
        _st_clamp_(w, self.wmin, self.wmax)
        d2 = w.log().pow(2).sum(-1)
        return _st_clamp_(d2, self.wmin)

    def _logdet(self, x):  # torch_batch.py:173-190 (PLogDet.forward; backward = inverse via the same factor)
        l = self.chol(x)
        return 2 * l.diagonal(dim1=-2, dim2=-1).abs().log().sum(-1), l

    def dist2(self, x, y):
        """Elementwise squared distance / divergence: spd.py:171-173, :183-189."""
        if self.stein:
            return _SteinFn.apply(x, y, self)
        a, _ = self.invchol(x)
        return self._norm_log2(_axat(a, y))

    def pdist2(self, x):
        """All a<b pairs, torch.triu_indices order: spd.py:175-181, :191-194."""
        i, j = torch.triu_indices(x.shape[0], x.shape[0], 1)
        if self.stein:
            return _SteinFn.apply(x[i], x[j], self)
        a, _ = self.invchol(x)
        return self._norm_log2(_axat(a[i], x[j]))

    # --- optimizer callees (spd.py:113-154,196-199) ---
    def egrad2rgrad(self, x, g):  # :134-135
        return _axat(x, _sym(g))

    def norm(self, x, u):  # :113-117, keepdim=True
        a, _ = self.invchol(x)
        return _axat(a, u).pow(2).sum((-2, -1), keepdim=True).sqrt()

    def exp(self, x, u):  # :137-144
        a, l = self.invchol(x)
        return _axat(l, _symapply(_axat(a, u), torch.exp))

    def retr(self, x, u):  # :146-154
        l = self.chol(x)
        w = torch.linalg.solve_triangular(l, u, upper=False)
        return _sym(x + u + 0.5 * torch.einsum('...ji,...jk->...ik', w, w))

    def log(self, x, y):  # :156-161
        a, l = self.invchol(x)
        return _axat(l, _symapply(_axat(a, y), torch.log))

    def transp(self, x, y, u):  # :196-199
        return u

    def projx(self, x):  # :126-132
        return _symapply(_sym(x), lambda w: w, self.wmin, self.wmax)

    def proju(self, x, u):  # :119-124
        return _sym(u)

    def inner(self, x, u, v):  # :100-106
        l = self.chol(x)
        xu = torch.cholesky_solve(u, l)
        xv = torch.cholesky_solve(v, l)
        return (xu @ xv).diagonal(dim1=-2, dim2=-1).sum(-1)

    def rand(self, n_points, ir=1e-1, dtype=torch.float64, generator=None):  # :201-208
        dim = self.n * (self.n + 1) // 2
        u = torch.randn(n_points, dim, dtype=dtype, generator=generator)
        u = u / u.norm(dim=-1, keepdim=True) * ir
        m = torch.zeros(n_points, self.n, self.n, dtype=dtype)
        iu = torch.triu_indices(self.n, self.n)
        m[:, iu[0], iu[1]] = u / math.sqrt(2)
        m = m + m.transpose(-2, -1)  # doubles the diagonal: u_ii/sqrt(2)*2 = sqrt(2) u_ii ... (from_vec, :75-80)
        d = torch.arange(self.n)
        m[:, d, d] = u[:, [int(k) for k in _diag_positions(self.n)]]
        eye = torch.eye(self.n, dtype=dtype).expand(n_points, -1, -1)
        return self.exp(eye, m)


def _diag_positions(n):
    pos, k = [], 0
    for i in range(n):
        pos.append(k)
        k += n - i
    return pos


class _SteinFn(torch.autograd.Function):
    """S(x,y) = logdet((x+y)/2) - (logdet x + logdet y)/2 with the hand-written backward
    of spd.py:246-295 / torch_batch.py:173-190: d/dx = ((x+y)/2)^-1 / 2 - x^-1 / 2, the
    inverses taken through the (possibly eps-perturbed) factor that `chol` returned."""

    @staticmethod
    def forward(ctx, x, y, man):
        z = 0.5 * (x + y)
        ldz, lz = man._logdet(z)
        ldx, lx = man._logdet(x)
        ldy, ly = man._logdet(y)
        ctx.save_for_backward(lz, lx, ly)
        s = ldz - 0.5 * (ldx + ldy)
        return s.clamp_(min=man.wmin)  # value-only in the reference (.data.clamp_)

    @staticmethod
    def backward(ctx, g):
        lz, lx, ly = ctx.saved_tensors
        eye = torch.eye(lz.shape[-1], dtype=lz.dtype).expand_as(lz)
        zi = torch.cholesky_solve(eye, lz)
        xi = torch.cholesky_solve(eye, lx)
        yi = torch.cholesky_solve(eye, ly)
        g = g.view(-1, 1, 1)
        return g * 0.5 * (zi - xi), g * 0.5 * (zi - yi), None


# ----------------------------------------------------------------------------
# Lorentz (manifolds/lorentz.py)
# ----------------------------------------------------------------------------
class _LDot(torch.autograd.Function):  # lorentz.py:101-118
    @staticmethod
    def forward(ctx, u, v):
        ctx.save_for_backward(u, v)
        uv = u * v
        uv[..., 0] *= -1
        return uv.sum(-1, keepdim=True)

    @staticmethod
    def backward(ctx, g):
        u, v = ctx.saved_tensors
        g = g.expand_as(u).clone()
        g[..., 0] *= -1
        return g * v, g * u


class _Acosh(torch.autograd.Function):  # lorentz.py:125-138
    @staticmethod
    def forward(ctx, x):
        z = torch.sqrt(x * x - 1)
        ctx.save_for_backward(z)
        return torch.log(x + z)

    @staticmethod
    def backward(ctx, g):
        z, = ctx.saved_tensors
        return g / z.clamp(min=EPS)


def ldot(u, v):
    return _LDot.apply(u, v)


class LorentzOracle:
    def __init__(self, n):
        self.n = n

    def dist2(self, x, y):  # lorentz.py:72-77
        d = -ldot(x, y).squeeze(-1)
        _st_clamp_(d, 1)
        dist = _Acosh.apply(d)
        _st_clamp_(dist, EPS)
        return dist.pow(2)

    def pdist2(self, x):  # base.py:59-63
        i, j = torch.triu_indices(x.shape[0], x.shape[0], 1)
        return self.dist2(x[i], x[j])

    def proju(self, x, u):  # :39-42
        return u + ldot(x, u) * x

    def egrad2rgrad(self, x, g):  # :52-57
        g = g.clone()
        g[..., 0] *= -1
        return self.proju(x, g)

    def norm(self, x, u):  # base.py:29-32 with inner = ldot(u,u) (:36-37), keepdim=True
        return ldot(u, u).clamp(min=EPS).sqrt()

    def exp(self, x, u):  # :59-62
        un = ldot(u, u).clamp(min=0).sqrt().clamp(min=EPS)
        return x * un.cosh() + un.sinh() * u / un

    retr = exp  # base.py:49-50

    def log(self, x, y):  # :64-70
        xy = ldot(x, y).clamp(max=-1)
        denom = torch.sqrt(xy * xy - 1).clamp(min=EPS)
        num = _Acosh.apply(-xy).clamp(min=EPS)
        return self.proju(x, num / denom * (y + xy * x))

    def transp(self, x, y, u):  # :79-82
        return u + ldot(u, y) / (1 - ldot(x, y)) * (x + y)

    def projx(self, x):  # :44-50
        x = x.clone()
        x[..., 0] = torch.sqrt(1 + x[..., 1:].pow(2).sum(-1))
        return x

    def inner(self, x, u, v):
        return ldot(u, v).squeeze(-1)

    def rand(self, n_points, ir=1e-2, dtype=torch.float64, generator=None):  # :84-86
        x = torch.empty(n_points, self.n, dtype=dtype).uniform_(-ir, ir, generator=generator)
        return self.projx(x)


# ----------------------------------------------------------------------------
# Sphere / Euclidean (manifolds/sphere.py, manifolds/euclidean.py)
# ----------------------------------------------------------------------------
class SphereOracle:
    def __init__(self, n):
        self.n = n

    def dist2(self, x, y):  # sphere.py:68-74
        s = (x * y).sum(-1)
        _st_clamp_(s, -1 + EPS**2, 1 - EPS**2)
        d = torch.acos(s)
        _st_clamp_(d, EPS)
        return d.pow(2)

    def pdist2(self, x):
        i, j = torch.triu_indices(x.shape[0], x.shape[0], 1)
        return self.dist2(x[i], x[j])

    def proju(self, x, u):  # :41-44
        return u - (x * u).sum(-1, keepdim=True) * x

    egrad2rgrad = proju  # base.py:42-43

    def norm(self, x, u):  # base.py:29-32
        return (u * u).sum(-1, keepdim=True).clamp(min=EPS).sqrt()

    def projx(self, x):  # :46-49
        return x / self.norm(None, x)

    def retr(self, x, u):  # :58-59
        return self.projx(x + u)

    def exp(self, x, u):  # :51-56
        nu = self.norm(None, u)
        e = x * torch.cos(nu) + u * torch.sin(nu) / nu
        return torch.where(nu > EPS, e, self.retr(x, u))

    def log(self, x, y):  # :61-66
        u = self.proju(x, y - x)
        d = self.dist2(x, y).sqrt().unsqueeze(-1)
        return torch.where(d > EPS, u * d / self.norm(None, u), u)

    def transp(self, x, y, u):  # base.py:65-66
        return self.proju(y, u)

    def inner(self, x, u, v):
        return (u * v).sum(-1)

    def rand_uniform(self, n_points, dtype=torch.float64, generator=None):  # :81-83
        return self.projx(torch.randn(n_points, self.n, dtype=dtype, generator=generator))


class EuclideanOracle:
    def __init__(self, n):
        self.n = n

    def dist2(self, x, y):  # euclidean.py:46-50 + base.py:29-32,56-57
        d = y - x
        return _st_clamp_((d * d).sum(-1), EPS)

    def pdist2(self, x):
        i, j = torch.triu_indices(x.shape[0], x.shape[0], 1)
        return self.dist2(x[i], x[j])

    def proju(self, x, u):
        return u

    egrad2rgrad = proju

    def norm(self, x, u):
        return (u * u).sum(-1, keepdim=True).clamp(min=EPS).sqrt()

    def exp(self, x, u):
        return x + u

    retr = exp

    def log(self, x, y):
        return y - x

    def transp(self, x, y, u):
        return u

    def projx(self, x):
        return x

    def inner(self, x, u, v):
        return (u * v).sum(-1)


# ----------------------------------------------------------------------------
# Universal: kappa-stereographic model (manifolds/universal.py over manifolds/impl/math.py)
# ----------------------------------------------------------------------------
MIN_NORM = 1e-15  # impl/math.py:15
BALL_EPS = {torch.float32: 4e-3, torch.float64: 1e-5}  # impl/math.py:16


class _Artanh(torch.autograd.Function):  # impl/math.py:25-40
    @staticmethod
    def forward(ctx, x):
        x = x.clamp(-1 + 1e-15, 1 - 1e-15)
        ctx.save_for_backward(x)
        xd = x.double()
        return ((1 + xd).log() - (1 - xd).log()).mul(0.5).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        x, = ctx.saved_tensors
        return g / (1 - x**2)


class UniversalOracle:
    """`c` is the tensor Universal.get_c() returns (shape (1,), may require grad)."""

    def __init__(self, n, c):
        self.n, self.c = n, c

    # impl/math.py:73-99 (scalar c: one branch)
    def _tan(self, x):
        return x.clamp(-15, 15).tanh() if self.c.item() > 0 else torch.tan(x)

    def _arctan(self, x):
        return _Artanh.apply(x) if self.c.item() > 0 else torch.atan(x)

    def _lambda(self, x, c=None):  # impl/math.py:187-190 (keepdim=True)
        c = self.c if c is None else c
        return 2 / (1 - c * x.pow(2).sum(-1, keepdim=True)).clamp_min(MIN_NORM)

    def _madd(self, x, y):  # impl/math.py:326-345
        c = self.c
        x2 = x.pow(2).sum(-1, keepdim=True)
        y2 = y.pow(2).sum(-1, keepdim=True)
        xy = (x * y).sum(-1, keepdim=True)
        num = (1 + 2 * c * xy + c * y2) * x + (1 - c * x2) * y
        den = 1 + 2 * c * xy + c**2 * x2 * y2
        return num / den.clamp_min(MIN_NORM)

    def projx(self, x):  # impl/math.py:142-156 (what Universal.projx(inplace=True) leaves in x)
        if self.c.item() <= 0:
            return x
        norm = x.norm(dim=-1, keepdim=True, p=2).clamp_min(MIN_NORM)
        maxnorm = (1 - BALL_EPS[x.dtype]) / self.c.abs().sqrt()
        return torch.where(norm > maxnorm, x / norm * maxnorm, x)

    def dist2(self, x, y):  # universal.py:76-81 (squared=True) + impl/math.py:567-572
        sc = self.c.abs()**0.5
        d = self._arctan(sc * self._madd(-x, y).norm(dim=-1, p=2)) * 2 / sc
        d = d.pow(2)
        return _st_clamp_(d, EPS)

    def dist(self, x, y):  # squared=False: the value clamp acts on d itself
        sc = self.c.abs()**0.5
        d = self._arctan(sc * self._madd(-x, y).norm(dim=-1, p=2)) * 2 / sc
        return _st_clamp_(d, EPS)

    def pdist2(self, x):  # base.py:59-63
        i, j = torch.triu_indices(x.shape[0], x.shape[0], 1)
        return self.dist2(x[i], x[j])

    def proju(self, x, u):  # universal.py:50-51
        return u

    def egrad2rgrad(self, x, g):  # impl/math.py:1452-1453
        return g / self._lambda(x)**2

    def norm(self, x, u):  # universal.py:44-48: math.norm is called WITHOUT c => c = 1.0; keepdim=True
        return self._lambda(x, c=1.0) * u.norm(dim=-1, keepdim=True, p=2)

    def inner(self, x, u, v):  # impl/math.py:225-228 with keepdim=False: the (N, 1) conformal factor (keepdim=True
        # is hard-coded there) broadcasts against the (N,) dot products into an (N, N) matrix -- kept as is
        return self._lambda(x)**2 * (u * v).sum(-1)

    def exp(self, x, u):  # universal.py:62-67, impl/math.py:720-727
        sc = self.c.abs()**0.5
        un = u.norm(dim=-1, p=2, keepdim=True).clamp_min(MIN_NORM)
        second = self._tan(sc / 2 * self._lambda(x) * un) * u / (sc * un)
        return self.projx(self._madd(x, second))

    def retr(self, x, u):  # universal.py:69-70
        return self.projx(x + u)

    def log(self, x, y):  # impl/math.py:835-841
        sub = self._madd(-x, y)
        sn = sub.norm(dim=-1, p=2, keepdim=True).clamp_min(MIN_NORM)
        sc = self.c.abs()**0.5
        return 2 / sc / self._lambda(x) * self._arctan(sc * sn) * sub / sn

    def transp(self, x, y, w):  # impl/math.py:1359-1362 with the gyration of :1282-1298 (u = y, v = -x)
        c = self.c
        u, v = y, -x
        u2 = u.pow(2).sum(-1, keepdim=True)
        v2 = v.pow(2).sum(-1, keepdim=True)
        uv = (u * v).sum(-1, keepdim=True)
        uw = (u * w).sum(-1, keepdim=True)
        vw = (v * w).sum(-1, keepdim=True)
        c2 = c**2
        a = -c2 * uw * v2 + c * vw + 2 * c2 * uv * vw
        b = -c2 * vw * u2 - c * uw
        d = 1 + 2 * c * uv + c2 * u2 * v2
        gyr = w + 2 * (a * u + b * v) / d.clamp_min(MIN_NORM)
        return gyr * self._lambda(x) / self._lambda(y)

    def rand(self, n_points, ir=1e-2, dtype=torch.float64, generator=None):  # universal.py:86-88
        x = torch.empty(n_points, self.n, dtype=dtype).uniform_(-ir, ir, generator=generator)
        return self.projx(x)


def universal_get_c(c_param, c_min=0.001, sign=None):  # universal.py:28-32
    if sign:
        return sign * (c_min + torch.nn.functional.softplus(c_param))
    return c_param.sign() * c_min + c_param


# ----------------------------------------------------------------------------
# Grassmann (manifolds/grassmann.py)
# ----------------------------------------------------------------------------
def singular_values_2x2_closed(a, eps=1e-8):  # fast.py:138-159
    p, q, r, s = a[..., 0, 0], a[..., 0, 1], a[..., 1, 0], a[..., 1, 1]
    s1 = p**2 + q**2 + r**2 + s**2
    s2 = (p**2 + q**2 - r**2 - s**2)**2 + 4 * (p * r + q * s)**2
    _st_clamp_(s2, eps)
    s2 = torch.sqrt(s2)
    big = 0.5 * (s1 + s2)
    _st_clamp_(big, eps)
    small = 0.5 * (s1 - s2)
    _st_clamp_(small, eps)
    return torch.stack([torch.sqrt(big), torch.sqrt(small)], dim=-1)


class GrassmannOracle:
    def __init__(self, n, p, retr='svd'):
        self.n, self.p, self.retr_kind = n, p, retr

    def dist2(self, x, y):  # grassmann.py:91-96
        a = torch.einsum('...ji,...jk->...ik', x, y)
        s = singular_values_2x2_closed(a) if self.p == 2 else torch.linalg.svd(a)[1]  # :27-30
        _st_clamp_(s, -1 + EPS**2, 1 - EPS**2)
        return s.acos().pow(2).sum(-1)

    def pdist2(self, x):
        i, j = torch.triu_indices(x.shape[0], x.shape[0], 1)
        return self.dist2(x[i], x[j])

    def proju(self, x, u):  # :49-53
        return u - x @ (x.transpose(-2, -1) @ u)

    egrad2rgrad = proju

    def norm(self, x, u):  # base.py:29-32 with :46-47
        return (u * u).sum((-2, -1), keepdim=True).clamp(min=EPS).sqrt()

    def projx(self, x):  # :55-61
        return torch.linalg.qr(x)[0]

    def exp(self, x, u):  # :63-69
        us, ss, vh = torch.linalg.svd(u, full_matrices=False)
        v = vh.transpose(-2, -1)
        lhs = x @ torch.einsum('...ij,...j,...kj->...ik', v, ss.cos(), v)
        return lhs + torch.einsum('...ij,...j,...kj->...ik', us, ss.sin(), v)

    def retr(self, x, u):  # :71-80
        if self.retr_kind == 'qr':
            return torch.linalg.qr(x + u)[0]
        us, _, vh = torch.linalg.svd(x + u, full_matrices=False)
        return us @ vh

    def log(self, x, y):  # :82-89
        ytx = y.transpose(-2, -1) @ x
        at = y.transpose(-2, -1) - ytx @ x.transpose(-2, -1)
        bt = torch.linalg.solve(ytx, at)
        us, ss, vh = torch.linalg.svd(bt.transpose(-2, -1), full_matrices=False)
        return torch.einsum('...ij,...j,...kj->...ik', us, ss.atan(), vh.transpose(-2, -1))

    def transp(self, x, y, u):  # base.py:65-66
        return self.proju(y, u)

    def inner(self, x, u, v):
        return (u * v).sum((-2, -1))

    def rand_uniform(self, n_points, dtype=torch.float64, generator=None):  # :105-107
        return self.projx(torch.randn(n_points, self.n, self.p, dtype=dtype, generator=generator))


# ----------------------------------------------------------------------------
# losses (objectives.py) and product embedding (modules.py:84-88)
# ----------------------------------------------------------------------------
def quotient_loss(g, m, alpha, epoch, inc_l1=True, inc_l2=True):  # objectives.py:24-33
    g = g * alpha
    loss = 0
    if inc_l1:
        loss = loss + (m / g - 1.0).abs().sum()
    if inc_l2:
        loss = loss + (g / (m + 1.0 / (epoch + 1)) - 1.0).abs().sum()
    return loss


def stress_loss(g, m):  # objectives.py:41-42
    return (m - g).pow(2).sum()


def _condensed_to_rows(theta):  # inference/stochastic_neighbors.py:13-18: n x (n-1) matrix of theta_ij, j != i
    n = math.ceil(math.sqrt(2 * theta.shape[0]))
    assert n * (n - 1) // 2 == theta.shape[0]
    tm = torch.ones(n, n - 1, dtype=torch.bool).triu()
    mat = theta.new_empty(n, n - 1)
    mat = mat.masked_scatter(tm, theta)              # row i, columns >= i  <- pairs (i, j > i)
    mat_t = mat.T.masked_scatter(~tm.T, theta)       # transposed fill     <- pairs (j < i, i)
    return mat_t.T, tm


def sne_log_partition(theta, ret_margs=False):  # inference/stochastic_neighbors.py:11-24
    mat, tm = _condensed_to_rows(theta)
    logz = torch.logsumexp(mat, dim=1).sum()
    if not ret_margs:
        return logz, None
    margs = torch.softmax(mat, dim=1)
    margs = margs.masked_select(tm) + margs.T.masked_select(~tm.T)
    return logz, margs


def kl_sne_loss(g, m, alpha, inclusive=True):  # objectives.py:61-74
    theta_x, theta_z = -alpha * g, -m
    if not inclusive:
        theta_x, theta_z = theta_z, theta_x
    a_x, margs_x = sne_log_partition(theta_x, ret_margs=True)
    a_z, _ = sne_log_partition(theta_z, ret_margs=False)
    return a_z - a_x - margs_x @ (theta_z - theta_x)


def pearsonr(x, y):  # metrics.py:13-17
    xm, ym = x - x.mean(), y - y.mean()
    return xm @ ym / (xm.norm() * ym.norm())


def average_distortion(mpdists, gpdists):  # metrics.py:46-56
    return torch.mean(torch.abs(mpdists - gpdists) / gpdists)


def validation_metrics(oracles, xs, scales, targets_sq_condensed):  # train.py:230-265 (the two default metrics)
    g = targets_sq_condensed.sqrt()
    m = product_dist2(oracles, xs, scales, lambda o, x: o.pdist2(x)).sqrt()
    return dict(pearsonr=pearsonr(m, g).item(), average_distortion=average_distortion(m, g).item())


def product_dist2(oracles, xs, scales, pair_fn):  # modules.py:84-88
    return sum(torch.nn.functional.softplus(s) * pair_fn(o, x) for o, x, s in zip(oracles, xs, scales))


def dataset_targets(hops_condensed, dtype):  # data/dataset.py:9-13 (condensed form; the dense matrix is its squareform)
    t = hops_condensed.to(dtype).pow(2)
    return t / t.max()


def batch_targets(dense, idx):  # data/dataset.py:19-27
    sub = dense[idx][:, idx]
    i, j = torch.triu_indices(len(idx), len(idx), 1)
    return sub[i, j]


# ----------------------------------------------------------------------------
# optimizers (optim/radam.py:43-98, optim/rsgd.py:40-82)
# ----------------------------------------------------------------------------
def radam_step(man, x, grad, state, lr, betas=(0.9, 0.999), max_grad_norm=None, exact=False, nc=False):
    """One RiemannianAdam update of parameter tensor x (returns new x; mutates `state` like the reference)."""
    if not state:
        state.update(step=1, exp_avg=torch.zeros_like(x), exp_avg_sq=torch.zeros_like(x))
    b1, b2 = betas
    step = state['step']
    retr = man.exp if exact else man.retr
    rg = man.egrad2rgrad(x, grad)
    gn = man.norm(x, rg)
    if max_grad_norm is not None:
        rg = rg * torch.clamp(max_grad_norm / gn, max=1.0)
    if nc:
        b2 = 1 - 1 / step
    m = state['exp_avg'] * b1 + (1 - b1) * rg
    v = state['exp_avg_sq'] * b2 + (1 - b2) * gn.pow(2)
    denom = v.sqrt() + EPS
    alpha = lr * (1 - b2**step)**0.5 / (1 - b1**step)
    direction = (denom / m).reciprocal() * (-alpha)
    new_x = retr(x, direction)
    state['exp_avg'] = man.transp(x, new_x, m)
    state['exp_avg_sq'] = v
    state['step'] = step + 1
    return new_x


def rsgd_step(man, x, grad, state, lr, momentum=0.0, dampening=0.0, max_grad_norm=None, exact=False):
    if not state and momentum > 0:
        state['momentum_buffer'] = grad.clone()  # rsgd.py:53-54 (Euclidean gradient!)
    retr = man.exp if exact else man.retr
    rg = man.egrad2rgrad(x, grad)
    if max_grad_norm is not None:
        gn = man.norm(x, rg)
        rg = rg * torch.clamp(max_grad_norm / gn, max=1.0)
    if momentum > 0:
        buf = state['momentum_buffer'] * momentum + (1 - dampening) * rg
        new_x = retr(x, -lr * buf)
        state['momentum_buffer'] = man.transp(x, new_x, buf)
        return new_x
    return retr(x, -lr * rg)
