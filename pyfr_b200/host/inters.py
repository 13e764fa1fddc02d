"""Interface (interior / inter-partition) kernel declarations.

Host-side counterpart of ``pyfr/solvers/base/inters.py:6-119``,
``pyfr/solvers/baseadvec/inters.py:9-58``,
``pyfr/solvers/baseadvecdiff/inters.py:7-66``,
``pyfr/solvers/euler/inters.py:12-49`` and
``pyfr/solvers/navstokes/inters.py:9-67``: builds the gather/scatter views
over the element buffers for each side of each interface, the scaled
normals, the memory-order permutation of interior interfaces, and declares
the ``intconu/intcflux/mpiconu/mpicflux`` kernels with the reference's
keyword arguments, and -- for boundaries -- the ``bcconu/bccflux`` kernels
with the per-type template arguments of ``pyfr/solvers/euler/inters.py:
50-110`` and ``pyfr/solvers/navstokes/inters.py:68-210``.
"""

import itertools as it
import math

import numpy as np


def _side_layout(side, elemap):
    """Start offset of every interface's run of flux points."""
    nfp = np.empty(len(side), dtype=np.int64)

    for etype, fidx, eidxs, where in side.foreach():
        nfp[where] = elemap[etype].nfacefpts[fidx]

    return np.concatenate(([0], np.cumsum(nfp)))


def side_view_maps(side, elemap, getter):
    """(matmap, rmap, cmap, rstridemap) arrays for a view over one side,
    one entry per flux point, in interface order."""
    start = _side_layout(side, elemap)
    n = int(start[-1])

    matmap = np.empty(n, dtype=np.int64)
    rmap = np.empty(n, dtype=np.int64)
    cmap = np.empty(n, dtype=np.int64)
    rsmap = np.ones(n, dtype=np.int64)

    for etype, fidx, eidxs, where in side.foreach():
        mid, rows, rstride = getattr(elemap[etype], getter)(eidxs, fidx)
        k = rows.shape[1]
        dst = (start[where][:, None] + np.arange(k)).ravel()

        matmap[dst] = mid
        rmap[dst] = rows.ravel()
        cmap[dst] = np.repeat(eidxs, k)
        if rstride is not None:
            rsmap[dst] = rstride

    return matmap, rmap, cmap, rsmap


def side_const(side, elemap, getter, ndims):
    start = _side_layout(side, elemap)
    out = np.empty((int(start[-1]), ndims))

    for etype, fidx, eidxs, where in side.foreach():
        vals = getattr(elemap[etype], getter)(eidxs, fidx)
        k = len(vals) // max(len(where), 1)
        dst = (start[where][:, None] + np.arange(k)).ravel()
        out[dst] = vals

    return out


class BaseInters:
    def __init__(self, be, lhs, elemap, cfg):
        self._be = be
        self.elemap = elemap
        self.cfg = cfg
        self.lhs = lhs

        e0 = next(iter(elemap.values()))
        self.ndims, self.nvars = e0.ndims, e0.nvars

        self.ninters = len(lhs)
        self.ninterfpts = sum(elemap[et].nfacefpts[fi]*len(ei)
                              for et, fi, ei in lhs.items())

        self._perm = Ellipsis
        self.c = cfg.items_as('constants', float)
        self.kernels = {}
        self.mpireqs = {}

    def _view(self, side, getter, vshape, xchg=False):
        m, r, c, rs = (a[self._perm]
                       for a in side_view_maps(side, self.elemap, getter))
        mk = self._be.xchg_view if xchg else self._be.view

        return mk(m, r, c, rs, vshape=vshape)

    def _scal_view(self, side, getter, **kw):
        return self._view(side, getter, (self.nvars,), **kw)

    def _vect_view(self, side, getter, **kw):
        return self._view(side, getter, (self.ndims, self.nvars), **kw)

    def _memory_order_perm(self, side):
        # Order interface points by the address of their scal_fpts entry.
        # The reference forms that key from a view of shape () -- for the
        # blocked layout elements 8 apart within a run of ``leaddim``
        # columns then tie, so five blocks interleave lane by lane (still
        # whole sectors per warp).  ``inters-order = address`` sorts by the
        # address in the nvars-wide layout instead: consecutive points are
        # the consecutive lanes of one flux-point row, which is what lets
        # ``conu-pairs`` move two points per 16-byte access.  The order of
        # the points of an interface does not affect any result.
        m, r, c, _ = side_view_maps(side, self.elemap,
                                    'get_scal_fpts_for_inters')
        if getattr(self._be, 'inters_order', 'reference') == 'address':
            key = self._be.view(m, r, c, vshape=(self.nvars,)).mapping.get()
            return np.argsort(key[0], kind='stable')
        return np.argsort(self._be.view(m, r, c, vshape=()).mapping.get()[0])

    def _pnorms(self, side):
        pn = side_const(side, self.elemap, 'get_pnorms_for_inters',
                        self.ndims)[self._perm]
        return self._be.const_matrix(np.atleast_2d(pn.T))

    def _rsolver_tplargs(self):
        be, cfg = self._be, self.cfg
        return dict(
            ndims=self.ndims, nvars=self.nvars, c=self.c,
            rsolver=cfg.get('solver-interfaces', 'riemann-solver'),
            p_min=cfg.getfloat('solver-interfaces', 'p-min',
                               5*be.fpdtype_eps)
        )


class IntInters(BaseInters):
    name = 'internal'

    def __init__(self, be, lhs, rhs, elemap, cfg):
        super().__init__(be, lhs, elemap, cfg)
        self.rhs = rhs

        self._perm = self._memory_order_perm(self._perm_side())
        self._pnorm_lhs = self._pnorms(lhs)

        g = 'get_scal_fpts_for_inters'
        self.scal_lhs = self._scal_view(lhs, g)
        self.scal_rhs = self._scal_view(rhs, g)

    def _perm_side(self):
        return self.lhs


class MPIInters(BaseInters):
    def __init__(self, be, lhs, rhsrank, rank, elemap, cfg):
        super().__init__(be, lhs, elemap, cfg)
        self.rhsrank, self.rank = rhsrank, rank
        self.name = f'p{rhsrank}'
        self._tags = it.count()

        self._pnorm_lhs = self._pnorms(lhs)

        self.scal_lhs = self._scal_view(lhs, 'get_scal_fpts_for_inters',
                                        xchg=True)
        self.scal_rhs = be.xchg_matrix_for_view(self.scal_lhs)

    def next_mpi_tag(self):
        return next(self._tags)


class EulerIntInters(IntInters):
    def __init__(self, *args):
        super().__init__(*args)
        be = self._be
        be.pointwise.register('pyfr.solvers.euler.kernels.intcflux')
        tplargs = self._rsolver_tplargs()

        self.kernels['comm_flux'] = lambda: be.kernel(
            'intcflux', tplargs=tplargs, dims=[self.ninterfpts],
            ul=self.scal_lhs, ur=self.scal_rhs, nl=self._pnorm_lhs
        )


class EulerMPIInters(MPIInters):
    def __init__(self, *args):
        super().__init__(*args)
        be = self._be
        be.pointwise.register('pyfr.solvers.euler.kernels.mpicflux')
        tplargs = self._rsolver_tplargs()

        self.kernels['comm_flux'] = lambda: be.kernel(
            'mpicflux', tplargs, dims=[self.ninterfpts],
            ul=self.scal_lhs, ur=self.scal_rhs, nl=self._pnorm_lhs
        )


def _ns_tplargs(inter):
    cfg = inter.cfg
    return inter._rsolver_tplargs() | dict(
        visc_corr=cfg.get('solver', 'viscosity-correction', 'none'),
        shock_capturing=cfg.get('solver', 'shock-capturing', 'none')
    )


class NavierStokesIntInters(IntInters):
    def __init__(self, *args):
        super().__init__(*args)
        be, lhs, rhs = self._be, self.lhs, self.rhs

        self._vect_lhs = self._vect_view(lhs, 'get_vect_fpts_for_inters')
        self._vect_rhs = self._vect_view(rhs, 'get_vect_fpts_for_inters')
        self._comm_lhs = self._scal_view(lhs, 'get_comm_fpts_for_inters')
        self._comm_rhs = self._scal_view(rhs, 'get_comm_fpts_for_inters')

        self.c |= self.cfg.items_as('solver-interfaces', float)
        tplargs = _ns_tplargs(self)

        be.pointwise.register('pyfr.solvers.navstokes.kernels.intconu')
        be.pointwise.register('pyfr.solvers.navstokes.kernels.intcflux')

        self.kernels['con_u'] = lambda: be.kernel(
            'intconu', tplargs=tplargs, dims=[self.ninterfpts],
            ulin=self.scal_lhs, urin=self.scal_rhs,
            ulout=self._comm_lhs, urout=self._comm_rhs
        )
        self.kernels['comm_flux'] = lambda: be.kernel(
            'intcflux', tplargs=tplargs, dims=[self.ninterfpts],
            ul=self.scal_lhs, ur=self.scal_rhs,
            gradul=self._vect_lhs, gradur=self._vect_rhs,
            artvisc=None, nl=self._pnorm_lhs
        )

    def _perm_side(self):
        beta = self.cfg.getfloat('solver-interfaces', 'ldg-beta')
        return self.lhs if beta != -0.5 else self.rhs


class NavierStokesMPIInters(MPIInters):
    def __init__(self, *args):
        super().__init__(*args)
        be, lhs, rank, rhsrank = self._be, self.lhs, self.rank, self.rhsrank

        self._vect_lhs = self._vect_view(lhs, 'get_vect_fpts_for_inters',
                                         xchg=True)
        self._vect_rhs = be.xchg_matrix_for_view(self._vect_lhs)
        self._comm_lhs = self._scal_view(lhs, 'get_comm_fpts_for_inters',
                                         xchg=True)
        self._comm_rhs = be.xchg_matrix_for_view(self._comm_lhs)

        self.c |= self.cfg.items_as('solver-interfaces', float)

        # One side of every partition boundary negates beta so the LDG
        # switch is consistent across it (reference baseadvecdiff
        # inters.py:47-58)
        if (rank + rhsrank) % 2:
            self.c['ldg-beta'] *= 1.0 if rank > rhsrank else -1.0
        else:
            self.c['ldg-beta'] *= 1.0 if rhsrank > rank else -1.0

        tplargs = _ns_tplargs(self)

        be.pointwise.register('pyfr.solvers.navstokes.kernels.mpiconu')
        be.pointwise.register('pyfr.solvers.navstokes.kernels.mpicflux')

        self.kernels['con_u'] = lambda: be.kernel(
            'mpiconu', tplargs=tplargs, dims=[self.ninterfpts],
            ulin=self.scal_lhs, urin=self.scal_rhs, ulout=self._comm_lhs
        )
        self.kernels['comm_flux'] = lambda: be.kernel(
            'mpicflux', tplargs=tplargs, dims=[self.ninterfpts],
            ul=self.scal_lhs, ur=self.scal_rhs,
            gradul=self._vect_lhs, gradur=self._vect_rhs,
            artvisc=None, nl=self._pnorm_lhs
        )


class BCInters(BaseInters):
    """Boundary interfaces: only a left-hand (interior) state exists; the
    ghost state is a function of it, the normal and the section's
    parameters (``pyfr/solvers/baseadvec/inters.py:52-119``)."""

    type = None
    cflux_state = None
    # (option names, defaults) turned into C expressions / numbers
    expr_opts, expr_defaults, eval_opts = (), {}, ()

    def __init__(self, be, lhs, elemap, cfgsect, cfg):
        super().__init__(be, lhs, elemap, cfg)
        self.cfgsect = cfgsect
        self.name = cfgsect.removeprefix('soln-bcs-')

        self._perm = self._memory_order_perm(lhs)
        self._pnorm_lhs = self._pnorms(lhs)
        self.scal_lhs = self._scal_view(lhs, 'get_scal_fpts_for_inters')

        self._external_args = {'t': 'scalar fpdtype_t'}
        self._external_vals = {}

        nd = self.ndims
        if self.eval_opts:
            from pyfr_b200.host.exprs import npeval
            cc = cfg.items_as('constants', float)
            for k in self.eval_opts:
                self.c[k] = float(npeval(cfg.getexpr(cfgsect, k), cc))

        opts = [o for o in self.expr_opts
                if o not in 'uvw' or 'uvw'.index(o) < nd]
        if opts:
            self.c |= self._exp_opts(opts, lhs, self.expr_defaults)

    def _exp_opts(self, opts, lhs, default={}):
        cfg, sect = self.cfg, self.cfgsect

        subs = cfg.items('constants')
        subs |= dict(x='ploc[0]', y='ploc[1]', z='ploc[2]')
        subs |= dict(abs='fabs', pi=str(math.pi))

        exprs = {}
        for k in opts:
            if k in default:
                exprs[k] = cfg.getexpr(sect, k, default[k], subs=subs)
            else:
                exprs[k] = cfg.getexpr(sect, k, subs=subs)

        if any('ploc' in ex for ex in exprs.values()) and \
           'ploc' not in self._external_args:
            pl = side_const(lhs, self.elemap, 'get_ploc_for_inters',
                            self.ndims)[self._perm]
            self._external_args['ploc'] = f'in fpdtype_t[{self.ndims}]'
            self._external_vals['ploc'] = self._be.const_matrix(
                np.atleast_2d(pl.T))

        return exprs

    def _tplargs(self):
        return self._rsolver_tplargs() | dict(bctype=self.type,
                                              ninters=self.ninters)


    def _extra_consts(self):
        pass


class EulerBCInters(BCInters):
    def __init__(self, *args):
        super().__init__(*args)
        be = self._be
        be.pointwise.register('pyfr.solvers.euler.kernels.bccflux')
        tplargs = self._tplargs()

        self.kernels['comm_flux'] = lambda: be.kernel(
            'bccflux', tplargs=tplargs, dims=[self.ninterfpts],
            extrns=self._external_args, ul=self.scal_lhs,
            nl=self._pnorm_lhs, **self._external_vals
        )


class NavierStokesBCInters(BCInters):
    def __init__(self, *args):
        super().__init__(*args)
        be, lhs = self._be, self.lhs
        self._extra_consts()

        self._vect_lhs = self._vect_view(lhs, 'get_vect_fpts_for_inters')
        self._comm_lhs = self._scal_view(lhs, 'get_comm_fpts_for_inters')
        self.c |= self.cfg.items_as('solver-interfaces', float)

        tplargs = self._tplargs() | dict(
            bccfluxstate=self.cflux_state,
            visc_corr=self.cfg.get('solver', 'viscosity-correction', 'none'),
            shock_capturing=self.cfg.get('solver', 'shock-capturing', 'none')
        )

        be.pointwise.register('pyfr.solvers.navstokes.kernels.bcconu')
        be.pointwise.register('pyfr.solvers.navstokes.kernels.bccflux')

        self.kernels['con_u'] = lambda: be.kernel(
            'bcconu', tplargs=tplargs, dims=[self.ninterfpts],
            extrns=self._external_args, ulin=self.scal_lhs,
            ulout=self._comm_lhs, nlin=self._pnorm_lhs,
            **self._external_vals
        )
        self.kernels['comm_flux'] = lambda: be.kernel(
            'bccflux', tplargs=tplargs, dims=[self.ninterfpts],
            extrns=self._external_args, ul=self.scal_lhs,
            gradul=self._vect_lhs, nl=self._pnorm_lhs, artvisc=None,
            **self._external_vals
        )


class NavierStokesSubInflowFtpttang(NavierStokesBCInters):
    """Total pressure / total temperature inflow with a prescribed flow
    angle (pyfr/solvers/navstokes/inters.py:170-196)."""

    type = 'sub-in-ftpttang'
    cflux_state = 'ghost'
    eval_opts = ('cpTt', 'pt', 'theta')

    def _extra_consts(self):
        gamma = self.cfg.getfloat('constants', 'gamma')
        self.c['Rdcp'] = (gamma - 1.0)/gamma

        theta = self.c.pop('theta')*np.pi/180.0
        vc = np.array([np.cos(theta), np.sin(theta), 1.0])

        if self.ndims == 3:
            from pyfr_b200.host.exprs import npeval
            cc = self.cfg.items_as('constants', float)
            phi = float(npeval(self.cfg.getexpr(self.cfgsect, 'phi'), cc))
            phi *= np.pi/180.0
            vc[:2] *= np.sin(phi)
            vc[2] *= np.cos(phi)

        self.c['vc'] = vc[:self.ndims]


def _bc(base, btype, cflux_state=None, expr_opts=(), expr_defaults={},
        eval_opts=()):
    return type(f'{base.__name__}_{btype}', (base,), dict(
        type=btype, cflux_state=cflux_state, expr_opts=expr_opts,
        expr_defaults=expr_defaults, eval_opts=eval_opts
    ))


_zero_vel = {'u': 0, 'v': 0, 'w': 0}

euler_bc_map = {c.type: c for c in [
    _bc(EulerBCInters, 'slp-adia-wall'),
    _bc(EulerBCInters, 'sup-out-fn'),
    _bc(EulerBCInters, 'sup-in-fa', expr_opts=('rho', 'p', 'u', 'v', 'w')),
    _bc(EulerBCInters, 'char-riem-inv',
        expr_opts=('rho', 'p', 'u', 'v', 'w')),
]}

navstokes_bc_map = {c.type: c for c in [
    _bc(NavierStokesBCInters, 'no-slp-adia-wall', 'ghost-imperm'),
    _bc(NavierStokesBCInters, 'no-slp-isot-wall', 'ghost-imperm',
        expr_opts=('u', 'v', 'w'), expr_defaults=_zero_vel,
        eval_opts=('cpTw',)),
    _bc(NavierStokesBCInters, 'slp-adia-wall', None),
    _bc(NavierStokesBCInters, 'char-riem-inv', 'ghost',
        expr_opts=('rho', 'p', 'u', 'v', 'w')),
    _bc(NavierStokesBCInters, 'sup-in-fa', 'ghost',
        expr_opts=('rho', 'p', 'u', 'v', 'w')),
    _bc(NavierStokesBCInters, 'sup-out-fn', 'ghost'),
    _bc(NavierStokesBCInters, 'sub-in-frv', 'ghost',
        expr_opts=('rho', 'u', 'v', 'w'), expr_defaults=_zero_vel),
    _bc(NavierStokesBCInters, 'sub-out-fp', 'ghost', expr_opts=('p',)),
    NavierStokesSubInflowFtpttang,
]}
