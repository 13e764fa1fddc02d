"""Reference-element numerics for tensor-product (quad, hex) elements.

Produces the constant flux-reconstruction operator matrices ``M0 .. M6``
and ``opmat('M4 - M6*M0')``-style expressions with the same definitions,
point orderings and clean-up rule as the reference (``pyfr/shapes.py:79-135``
operator definitions, ``:382-482`` quad/hex faces; ``pyfr/nputil.py:23-57``
``clean``; ``pyfr/polys.py`` orthonormal Legendre bases).  The formulation
is our own: every operator is assembled from 1-D Lagrange/Legendre factors,
so no (p+1)^3-sized Vandermonde matrix is ever inverted.
"""

from functools import cached_property
import itertools as it
import os
import re

import numpy as np
from numpy.polynomial import legendre as npleg


# -- 1-D building blocks ---------------------------------------------------
def gauss_legendre(n):
    """n-point Gauss-Legendre nodes and weights, polished in extended
    precision so they round to the same doubles as a 36-digit table."""
    x, _ = npleg.leggauss(n)
    x = x.astype(np.longdouble)
    c = np.zeros(n + 1, dtype=np.longdouble)
    c[n] = 1

    for _ in range(3):
        x -= npleg.legval(x, c) / npleg.legval(x, npleg.legder(c))

    w = 2 / ((1 - x*x)*npleg.legval(x, npleg.legder(c))**2)
    x = 0.5*(x - x[::-1])

    return x.astype(float), w.astype(float)


def gauss_legendre_lobatto(n):
    c = np.zeros(n, dtype=np.longdouble)
    c[n - 1] = 1
    dc = npleg.legder(c)

    xi = np.sort(npleg.legroots(dc.astype(float))).astype(np.longdouble)
    for _ in range(3):
        xi -= npleg.legval(xi, dc) / npleg.legval(xi, npleg.legder(dc))

    x = np.concatenate(([-1], xi, [1])).astype(np.longdouble)
    w = 2 / (n*(n - 1)*npleg.legval(x, c)**2)
    x = 0.5*(x - x[::-1])

    return x.astype(float), w.astype(float)


_line_rules = {
    'gauss-legendre': gauss_legendre,
    'gauss-legendre-lobatto': gauss_legendre_lobatto
}


def orthonormal_legendre(order, x):
    """psi_i(x) = sqrt(i + 1/2) P_i(x), i = 0..order; shape (len(x), order+1)"""
    x = np.atleast_1d(np.asarray(x, dtype=float))
    v = npleg.legvander(x, order)
    return v*np.sqrt(np.arange(order + 1) + 0.5)


def orthonormal_legendre_deriv(order, x):
    x = np.atleast_1d(np.asarray(x, dtype=float))
    out = np.zeros((len(x), order + 1))

    for i in range(1, order + 1):
        c = np.zeros(i + 1)
        c[i] = 1
        out[:, i] = npleg.legval(x, npleg.legder(c))*np.sqrt(i + 0.5)

    return out


class Lagrange1D:
    """Nodal (Lagrange) basis through a 1-D node set."""

    def __init__(self, nodes):
        self.nodes = np.asarray(nodes, dtype=float)
        self.order = len(self.nodes) - 1
        self._ivdm = np.linalg.inv(orthonormal_legendre(self.order,
                                                        self.nodes))

    def at(self, x):
        # (len(x), nnodes): L_j(x_i)
        return orthonormal_legendre(self.order, x) @ self._ivdm

    def deriv_at(self, x):
        return orthonormal_legendre_deriv(self.order, x) @ self._ivdm


def clean(arr, tol=1e-10):
    """Flush |a| < tol to zero and snap magnitudes that agree to within
    ``tol`` onto their median, so that baked-constant kernels see a
    minimal set of unique values (rule of pyfr/nputil.py:23-57)."""
    arr = np.array(arr, dtype=float)
    arr[np.abs(arr) < tol] = 0

    if arr.size > 1:
        mag = np.abs(arr).ravel()
        order = np.argsort(mag)
        srt = mag[order].tolist()

        # A run continues while values stay close to the run's first entry
        start, ref, n = 0, srt[0], len(srt)
        for j in range(1, n + 1):
            if j == n or abs(srt[j] - ref) > 0.1*tol + tol*abs(ref):
                if j - start > 1:
                    mag[order[start:j]] = np.median(srt[start:j])
                if j < n:
                    start, ref = j, srt[j]

        arr = np.copysign(mag, arr.ravel()).reshape(arr.shape)

    return arr


# -- tensor-product shapes -------------------------------------------------
class TensorShape:
    name = None
    ndims = None
    faces = None          # (face kind, projection, unit normal)
    jac_exprs = None
    interp_expr = None

    def __init__(self, nspts, cfg):
        self.nspts = nspts
        self.cfg = cfg
        self.order = cfg.getint('solver', 'order')

        aa = cfg.get('solver', 'anti-alias', 'none')
        self.antialias = {s.strip() for s in aa.split(',')} - {'none'}
        if self.antialias - {'flux', 'surf-flux'}:
            raise ValueError('Invalid anti-alias options')

        n = self.order + 1
        urule = cfg.get(f'solver-elements-{self.name}', 'soln-pts')
        self._u1d, self._uw1d = _line_rules[urule](n)
        self._ubasis1d = Lagrange1D(self._u1d)

        fkind = 'line' if self.ndims == 2 else 'quad'
        frule = cfg.get(f'solver-interfaces-{fkind}', 'flux-pts')
        self._f1d, self._fw1d = _line_rules[frule](n)

        if nspts:
            self.nsptsord = round(nspts**(1/self.ndims)) - 1
            if (self.nsptsord + 1)**self.ndims != nspts:
                raise ValueError('Invalid number of shape points')

    # Point sets; first coordinate varies fastest throughout
    @classmethod
    def std_ele(cls, sptord):
        p1 = np.linspace(-1, 1, sptord + 1)
        return np.array([p[::-1] for p in it.product(p1, repeat=cls.ndims)])

    @staticmethod
    def _tensor_pts(x1d, ndims):
        return np.array([p[::-1] for p in it.product(x1d, repeat=ndims)])

    @cached_property
    def upts(self):
        return self._tensor_pts(self._u1d, self.ndims)

    @cached_property
    def upts_wts(self):
        w = self._uw1d
        for _ in range(self.ndims - 1):
            w = np.multiply.outer(self._uw1d, w)
        return w.ravel()

    @cached_property
    def nupts(self):
        return (self.order + 1)**self.ndims

    # Flux anti-aliasing: element quadrature rule (reference shapes.py:
    # 144-192 -- Gauss-Legendre, (order + 2) points a direction unless
    # quad-deg / quad-npts say otherwise)
    @cached_property
    def _q1d(self):
        sect = f'solver-elements-{self.name}'
        rule = self.cfg.get(sect, 'quad-pts', 'gauss-legendre')

        if self.cfg.hasopt(sect, 'quad-deg'):
            # By degree alone the reference picks the smallest tabulated
            # (non-tensor) rule; only named tensor rules are built here
            if not self.cfg.hasopt(sect, 'quad-pts'):
                raise NotImplementedError('quad-deg needs an explicit '
                                          'tensor-product quad-pts rule')
            n = self.cfg.getint(sect, 'quad-deg')//2 + 1
        elif self.cfg.hasopt(sect, 'quad-npts'):
            n = round(self.cfg.getint(sect, 'quad-npts')**(1/self.ndims))
        else:
            n = self.order + 2

        return _line_rules[rule](n)

    @cached_property
    def qpts(self):
        return self._tensor_pts(self._q1d[0], self.ndims)

    @cached_property
    def qpts_wts(self):
        w1 = self._q1d[1]
        w = w1
        for _ in range(self.ndims - 1):
            w = np.multiply.outer(w1, w)
        return w.ravel()

    @property
    def nqpts(self):
        return len(self.qpts)

    # Surface-flux anti-aliasing: the flux points become the points of a
    # face quadrature rule and the common flux is L2-projected back onto
    # the face polynomials inside M3 (reference shapes.py:104-113,182-211)
    @cached_property
    def _fq1d(self):
        fkind = 'line' if self.ndims == 2 else 'quad'
        sect = f'solver-interfaces-{fkind}'
        rule = self.cfg.get(sect, 'quad-pts', 'gauss-legendre')

        if self.cfg.hasopt(sect, 'quad-deg'):
            if fkind != 'line' and not self.cfg.hasopt(sect, 'quad-pts'):
                raise NotImplementedError('quad-deg needs an explicit '
                                          'tensor-product quad-pts rule')
            n = self.cfg.getint(sect, 'quad-deg')//2 + 1
        elif self.cfg.hasopt(sect, 'quad-npts'):
            n = round(self.cfg.getint(sect, 'quad-npts')
                      **(1/(self.ndims - 1)))
        else:
            n = self.order + 2

        return _line_rules[rule](n)

    @cached_property
    def _face_proj(self):
        """``(nfp, nq)``: values at the face quadrature points -> nodal
        values at the regular flux points of the degree-p L2 projection."""
        qx, qw = self._fq1d
        psi_f = orthonormal_legendre(self.order, self._f1d)      # (nfp, nb)
        psi_q = orthonormal_legendre(self.order, qx)             # (nq, nb)
        p1 = psi_f @ (psi_q*qw[:, None]).T

        p = p1
        for _ in range(self.ndims - 2):
            p = np.kron(p1, p)
        return p

    @cached_property
    def _face_ref_pts(self):
        # Flux points and quadrature weights on the reference face
        pts = self._tensor_pts(self._f1d, self.ndims - 1)
        w = self._fw1d
        for _ in range(self.ndims - 2):
            w = np.multiply.outer(self._fw1d, w)
        return pts, w.ravel()

    @cached_property
    def fpts(self):
        if 'surf-flux' in self.antialias:
            fp = self._tensor_pts(self._fq1d[0], self.ndims - 1)
        else:
            fp, _ = self._face_ref_pts
        out = []

        for kind, proj, norm in self.faces:
            cols = np.broadcast_arrays(*proj(*fp.T))
            out.append(np.stack(cols, axis=1).astype(float))

        return np.vstack(out)

    @cached_property
    def nfacefpts(self):
        n = (len(self._fq1d[0]) if 'surf-flux' in self.antialias else
             self.order + 1)
        return [n**(self.ndims - 1)]*len(self.faces)

    @property
    def nfpts(self):
        return sum(self.nfacefpts)

    @cached_property
    def facefpts(self):
        off = np.cumsum([0] + self.nfacefpts)
        return [list(range(off[i], off[i + 1])) for i in range(len(off) - 1)]

    @cached_property
    def norm_fpts(self):
        return np.vstack([[norm]*n for (_, _, norm), n
                          in zip(self.faces, self.nfacefpts)]).astype(float)

    @cached_property
    def linspts(self):
        return self.std_ele(1)

    @cached_property
    def spts(self):
        return self.std_ele(self.nsptsord)

    @cached_property
    def mpts(self):
        return self.std_ele(max(self.order, 1))

    @cached_property
    def nmpts(self):
        return len(self.mpts)

    # Nodal bases evaluated at arbitrary points, assembled from 1-D factors
    @staticmethod
    def _tensor_eval(b1d, pts, deriv=None):
        pts = np.atleast_2d(pts)
        nd = pts.shape[1]
        fac = [b1d.deriv_at(pts[:, d]) if d == deriv else b1d.at(pts[:, d])
               for d in range(nd)]

        # out[p, i + n*j (+ n^2*k)]
        out = fac[0]
        for f in fac[1:]:
            out = (f[:, :, None]*out[:, None, :]).reshape(len(pts), -1)

        return out

    def ubasis_at(self, pts, clean_=True):
        m = self._tensor_eval(self._ubasis1d, pts)
        return clean(m) if clean_ else m

    def ubasis_deriv_at(self, pts, d):
        return self._tensor_eval(self._ubasis1d, pts, deriv=d)

    @cached_property
    def _sbasis1d(self):
        return Lagrange1D(np.linspace(-1, 1, self.nsptsord + 1))

    @cached_property
    def _mbasis1d(self):
        return Lagrange1D(np.linspace(-1, 1, max(self.order, 1) + 1))

    def sbasis_at(self, pts):
        return clean(self._tensor_eval(self._sbasis1d, pts))

    def mbasis_at(self, pts):
        return clean(self._tensor_eval(self._mbasis1d, pts))

    def mbasis_deriv_at(self, pts, d):
        return clean(self._tensor_eval(self._mbasis1d, pts, deriv=d))

    def _ortho_at(self, pts):
        # Orthonormal tensor Legendre basis, same index order as the nodes
        pts = np.atleast_2d(pts)
        fac = [orthonormal_legendre(self.order, pts[:, d])
               for d in range(self.ndims)]

        out = fac[0]
        for f in fac[1:]:
            out = (f[:, :, None]*out[:, None, :]).reshape(len(pts), -1)

        return out

    # Operator matrices (definitions: reference shapes.py:92-135)
    @cached_property
    def m0(self):
        return self.ubasis_at(self.fpts)

    @cached_property
    def m1(self):
        d = [self.ubasis_deriv_at(self.upts, i) for i in range(self.ndims)]
        return clean(np.hstack(d))

    @cached_property
    def m2(self):
        m = self.norm_fpts[..., None]*self.m0[:, None, :]
        return m.reshape(self.nfpts, -1)

    @cached_property
    def m3(self):
        # Divergence of the (DG) correction functions: for every face,
        # project the face Lagrange basis onto the orthonormal volume
        # basis with the face quadrature, then evaluate at the soln pts
        fp, fw = self._face_ref_pts
        fb1d = Lagrange1D(self._f1d)
        qx, qw = gauss_legendre(self.order + 1)
        qpts = self._tensor_pts(qx, self.ndims - 1)
        qwts = qw
        for _ in range(self.ndims - 2):
            qwts = np.multiply.outer(qw, qwts)
        qwts = qwts.ravel()

        lface = self._tensor_eval(fb1d, qpts)           # (nq, nfp)
        psi_u = self._ortho_at(self.upts)               # (nupts, nb)

        blocks = []
        for kind, proj, norm in self.faces:
            cols = np.broadcast_arrays(*proj(*qpts.T))
            vq = np.stack(cols, axis=1).astype(float)
            psi_q = self._ortho_at(vq)                  # (nq, nb)
            s = np.einsum('q,qf,qb->fb', qwts, lface, psi_q)
            blocks.append(psi_u @ s.T)                  # (nupts, nfp)

        if 'surf-flux' in self.antialias:
            blocks = [clean(b) @ self._face_proj for b in blocks]

        return clean(np.hstack(blocks))

    @cached_property
    def m4(self):
        m = self.m1.reshape(self.nupts, -1, self.nupts).swapaxes(0, 1)
        return m.reshape(-1, self.nupts)

    @cached_property
    def m6(self):
        m = self.norm_fpts.T[:, None, :]*self.m3
        return m.reshape(-1, self.nfpts)

    @cached_property
    def m7(self):
        # solution points -> quadrature points
        return self.ubasis_at(self.qpts)

    @cached_property
    def m8(self):
        # L2 projection from the quadrature points back onto the nodal
        # basis (reference shapes.py:18-19, 130-131)
        psi_u = self._ortho_at(self.upts)
        psi_q = self._ortho_at(self.qpts)
        return clean(psi_u @ (psi_q*self.qpts_wts[:, None]).T)

    @property
    def m9(self):
        nd, (a, b) = self.ndims, self.m8.shape
        m = np.zeros((nd*a, nd*b))
        for d in range(nd):
            m[d*a:(d + 1)*a, d*b:(d + 1)*b] = self.m8
        return m

    def opmat(self, expr):
        expr = expr.lower().replace('*', '@')

        if not re.match(r'[m0-9\-+@() ]+$', expr):
            raise ValueError('Invalid operator matrix expression')

        mats = {m: getattr(self, m) for m in re.findall(r'm\d+', expr)}
        return clean(eval(expr, {'__builtins__': None}, mats))

    @cached_property
    def fpts_in_upts(self):
        return bool(self.order > 0 and
                    np.all(np.count_nonzero(self.m0, axis=1) == 1))

    @cached_property
    def fpts_map_upts(self):
        if not self.fpts_in_upts:
            raise ValueError('Flux points not subset of solution points')

        return np.argwhere(np.abs(self.m0 - 1) <= 1e-8)[:, 1]


def _lin_jac_exprs(ndims):
    """C expressions for the Jacobian of the multilinear map through the
    2^ndims vertices ``V`` at reference point ``x``: d x_phys[i] / d xi[d]."""
    verts = list(it.product((-1, 1), repeat=ndims))
    verts = [v[::-1] for v in verts]

    rows = []
    for d in range(ndims):
        row = []
        for i in range(ndims):
            terms = []
            for n, v in enumerate(verts):
                f = [f'({"-" if v[e] < 0 else "+"}1)' if e == d else
                     f'(1 {"-" if v[e] < 0 else "+"} x[{e}])'
                     for e in range(ndims)]
                terms.append('*'.join(f) + f'*V[{n}][{i}]')
            row.append('(' + ' + '.join(terms) + f')/{2**ndims}')
        rows.append(row)

    return rows


class QuadShape(TensorShape):
    name = 'quad'
    ndims = 2

    faces = [
        ('line', lambda s: (s, -1), (0, -1)),
        ('line', lambda s: (1, s), (1, 0)),
        ('line', lambda s: (s, 1), (0, 1)),
        ('line', lambda s: (-1, s), (-1, 0)),
    ]

    jac_exprs = _lin_jac_exprs(2)


class HexShape(TensorShape):
    name = 'hex'
    ndims = 3

    faces = [
        ('quad', lambda s, t: (s, t, -1), (0, 0, -1)),
        ('quad', lambda s, t: (s, -1, t), (0, -1, 0)),
        ('quad', lambda s, t: (1, s, t), (1, 0, 0)),
        ('quad', lambda s, t: (s, 1, t), (0, 1, 0)),
        ('quad', lambda s, t: (-1, s, t), (-1, 0, 0)),
        ('quad', lambda s, t: (s, t, 1), (0, 0, 1)),
    ]

    jac_exprs = _lin_jac_exprs(3)


class TabulatedShape:
    """Simplex, prism and pyramid elements from tabulated data.

    Their orthonormal bases and point sets (Williams-Shunn, Shunn-Ham, ...)
    are not constructed here: ``data/tabshapes.npz`` holds, for the point
    sets of BASELINE.json configs[3] and orders 1-3, the operator matrices,
    point sets, weights, normals and nodal-basis evaluations produced by
    the reference's ``pyfr/shapes.py`` (generator:
    ``tests/golden/make_golden.py --shapes``).  Linear (straight-sided)
    elements only; no anti-aliasing."""

    name = ndims = None
    _data = None
    _rules = {
        'tri': 'williams-shunn', 'tet': 'shunn-ham',
        'pri': 'williams-shunn~gauss-legendre', 'pyr': 'gauss-legendre',
    }

    def __init__(self, nspts, cfg):
        if TabulatedShape._data is None:
            path = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                'data', 'tabshapes.npz')
            TabulatedShape._data = dict(np.load(path))

        self.nspts, self.cfg = nspts, cfg
        self.order = cfg.getint('solver', 'order')
        self.antialias = set()

        aa = cfg.get('solver', 'anti-alias', 'none')
        rule = cfg.get(f'solver-elements-{self.name}', 'soln-pts')
        self._pre = pre = f'{self.name}_p{self.order}_'

        if aa != 'none' or rule != self._rules[self.name] or \
           pre + 'upts' not in self._data:
            raise NotImplementedError(
                f'{self.name} elements are tabulated for orders 1-3, the '
                f'{self._rules[self.name]} points and no anti-aliasing'
            )
        if nspts != len(self.linspts):
            raise NotImplementedError('curved simplex/prism/pyramid elements')

        self.nsptsord = 1

    def _get(self, key):
        return self._data[self._pre + key]

    upts = property(lambda self: self._get('upts'))
    upts_wts = property(lambda self: self._get('upts_wts'))
    fpts = property(lambda self: self._get('fpts'))
    mpts = property(lambda self: self._get('mpts'))
    linspts = property(lambda self: self._get('linspts'))
    spts = property(lambda self: self._get('linspts'))
    norm_fpts = property(lambda self: self._get('norm_fpts'))
    nupts = property(lambda self: len(self._get('upts')))
    nfpts = property(lambda self: len(self._get('fpts')))
    nmpts = property(lambda self: len(self._get('mpts')))
    nqpts = qpts = None
    fpts_in_upts = False

    @cached_property
    def nfacefpts(self):
        return self._get('nfacefpts').tolist()

    @cached_property
    def facefpts(self):
        flat = self._get('facefpts').tolist()
        off = np.cumsum([0] + self.nfacefpts)
        return [flat[a:b] for a, b in zip(off, off[1:])]

    @cached_property
    def faceverts(self):
        flat = self._get('faceverts').tolist()
        off = np.cumsum([0] + self._get('nfaceverts').tolist())
        return [flat[a:b] for a, b in zip(off, off[1:])]

    @cached_property
    def jac_exprs(self):
        return [[str(e) for e in row] for row in self._get('jac_exprs')]

    def _named(self, pts):
        pts = np.atleast_2d(np.asarray(pts, dtype=float))
        for n in ('upts', 'fpts', 'mpts', 'linspts'):
            ref = self._get(n)
            if ref.shape == pts.shape and np.array_equal(ref, pts):
                return n
        raise NotImplementedError('basis evaluation away from the tabulated '
                                  'point sets')

    def sbasis_at(self, pts):
        return self._get(f'sbasis@{self._named(pts)}')

    def mbasis_at(self, pts):
        return self._get(f'mbasis@{self._named(pts)}')

    def mbasis_deriv_at(self, pts, d):
        if self._named(pts) != 'mpts':
            raise NotImplementedError('metric basis derivative away from the '
                                      'metric points')
        return self._get('mbasis_deriv@mpts')[d]

    def opmat(self, expr):
        expr = expr.lower().replace('*', '@')

        if not re.match(r'[m0-9\-+@() ]+$', expr):
            raise ValueError('Invalid operator matrix expression')

        mats = {m: self._get(m) for m in re.findall(r'm\d+', expr)}
        return clean(eval(expr, {'__builtins__': None}, mats))


def _tabulated(name_, ndims_):
    return type(f'{name_.title()}Shape', (TabulatedShape,),
                dict(name=name_, ndims=ndims_))


shape_map = {'quad': QuadShape, 'hex': HexShape,
             'tri': _tabulated('tri', 2), 'tet': _tabulated('tet', 3),
             'pri': _tabulated('pri', 3), 'pyr': _tabulated('pyr', 3)}
