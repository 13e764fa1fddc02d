"""TEST ORACLE -- a NumPy execution backend with the reference's API.

Test infrastructure only (see oracle/physics.py for who may import it).
``make_backend(base)`` builds the backend classes on top of either the
reference's own ``pyfr.backends.base`` package (when /root/reference is
importable; used to pin this repo's host mirror against the reference's
host code) or this repo's ``pyfr_b200.base`` mirror (everywhere else).

It honours the real storage layout -- matrices live at their byte offset
inside extents and kernels address them with the argument-dereference
rules of ``pyfr/backends/base/generator.py:171-276`` -- so view
``mapping``/``rstrides`` arrays are exercised exactly as a device backend
would exercise them.  Kernel arithmetic comes from oracle/physics.py.
"""

import numpy as np
from numpy.lib.stride_tricks import as_strided

from oracle import physics as ph


def _mat3(m, padto=1):
    """Strided (nblocks, nrow, width) window onto a matrix or slice."""
    it = m.itemsize
    flat = m.basedata.view(m.dtype)
    blocked = m.backend.blocks and not ({'xchg', 'noblock'} & set(m.tags))
    width = m.leaddim if blocked else -(-m.ncol // padto)*padto
    nblocks = m.nblocks if blocked else 1

    return as_strided(flat[m.offset // it:],
                      shape=(nblocks, m.nrow, width),
                      strides=(m.blocksz*it, m.leaddim*it, it))


def _stacked(m, nvar, nlead=None):
    """View with the variable (and leading stack) axes in front:
    ``[nlead,] nvar, nblocks, npts, nchunks, k``."""
    k = m.backend.soasz
    a = _mat3(m, padto=nvar*k)
    nb, nrow, w = a.shape
    a = a.reshape(nb, nrow, w // (nvar*k), nvar, k)

    if nlead is None:
        return a.transpose(3, 0, 1, 2, 4)
    else:
        a = a.reshape(nb, nlead, nrow // nlead, w // (nvar*k), nvar, k)
        return a.transpose(1, 4, 0, 2, 3, 5)


def _plain(m):
    """(nblocks, nrow, nchunks, k) window on a plain 2-D matrix."""
    k = m.backend.soasz
    a = _mat3(m, padto=k)
    return a.reshape(a.shape[0], a.shape[1], a.shape[2] // k, k)


class _ViewRef:
    """Gather/scatter through a view: v[r][c] <-> base[map + rstr*r + k*c]"""

    def __init__(self, view):
        if hasattr(view, 'view'):
            view = view.view

        self.flat = view.basedata.view(view.refdtype)
        self.map = view.mapping.get()[0].astype(np.int64)
        self.k = view.mapping.backend.soasz
        self.rstr = (view.rstrides.get()[0].astype(np.int64)
                     if view.rstrides is not None else None)

    def load(self, c, r=None):
        ix = self.map + self.k*c
        if r is not None:
            ix = ix + self.rstr*r
        return self.flat[ix]

    def store(self, val, c, r=None):
        ix = self.map + self.k*c
        if r is not None:
            ix = ix + self.rstr*r
        self.flat[ix] = val


def make_backend(base, name='oracle-numpy'):
    class NPMatrixBase(base.MatrixBase):
        def onalloc(self, basedata, offset):
            self.basedata, self.offset = basedata, offset
            self.data = basedata[offset:offset + self.nbytes].view(self.dtype)

            if self._initval is not None:
                self._set(self._initval)

            del self._initval

        def _get(self):
            shp = (self.nblocks, self.nrow, self.leaddim)
            return np.array(self._unpack(self.data.reshape(shp)))

        def _set(self, ary):
            self.data[:] = np.asarray(self._pack(ary)).reshape(-1)

    class NPMatrix(NPMatrixBase, base.Matrix): pass
    class NPConstMatrix(NPMatrixBase, base.ConstMatrix): pass
    class NPMatrixSlice(base.MatrixSlice): pass
    class NPView(base.View): pass
    class NPXchgView(base.XchgView): pass

    class NPXchgMatrix(NPMatrix, base.XchgMatrix):
        def recvreq(self, comm, pid, tag):
            return comm.recv_init(self, pid, tag)

        def sendreq(self, comm, pid, tag):
            return comm.send_init(self, pid, tag)

    class NPKernel(base.Kernel):
        def __init__(self, fn, rtnames=()):
            super().__init__()
            self._fn = fn
            self._rt = {}

            if rtnames:
                self.rtnames = tuple(rtnames)
                self.bind = self._bind

        def _bind(self, **kw):
            self._rt.update(kw)

        def add_to_graph(self, graph, deps):
            graph.program.append(('kernel', self))
            return self

        def run(self, *args):
            with np.errstate(all='ignore'):
                self._fn(**self._rt)

    def _meta(cls):
        class Meta(cls):
            def add_to_graph(self, graph, deps):
                graph.program.append(('kernel', self))
                return self
        return Meta

    ordered_cls = getattr(base, 'OrderedMetaKernel', None) or \
        base.BaseOrderedMetaKernel
    unordered_cls = getattr(base, 'UnorderedMetaKernel', None) or \
        base.BaseUnorderedMetaKernel

    class NPGraph(base.Graph):
        def __init__(self, backend):
            self.program = []
            super().__init__(backend)

        def _add_mpi_req(self, req, deps):
            super()._add_mpi_req(req, deps)
            self.program.append(('send' if deps else 'recv', req))

        def run(self, *args):
            for what, obj in self.program:
                if what == 'kernel':
                    obj.run()
                else:
                    obj.start()

        def get_wait_times(self):
            return []

    # -- providers -------------------------------------------------------
    class BlasProvider:
        def __init__(self, backend):
            self.backend = backend

        def mul(self, a, b, out, alpha=1.0, beta=0.0):
            if a.nrow != out.nrow or a.ncol != b.nrow or b.ncol != out.ncol:
                raise ValueError('Incompatible matrices for out = a*b')

            A = a.get()
            ext = self.backend.extended_mul

            # Kept so that a test can form the magnitude of the terms an
            # output is summed from (running-error bound, tests/util.py)
            self.backend.mul_log.append((A, b, out, alpha, beta))

            def run():
                B, C = _mat3(b), _mat3(out)

                if ext:
                    # Extended-precision accumulation: measures how much of
                    # a discrepancy is plain fp64 summation-order noise
                    r = alpha*np.einsum('mk,bkn->bmn',
                                        A.astype(np.longdouble),
                                        B.astype(np.longdouble))
                    C[:] = (r + beta*C if beta else r).astype(C.dtype)
                else:
                    r = alpha*np.einsum('mk,bkn->bmn', A, B)
                    C[:] = r + beta*C if beta else r

            return NPKernel(run)

        def copy(self, dst, src):
            if dst.traits != src.traits:
                raise ValueError('Incompatible matrix types')

            def run():
                _mat3(dst)[:] = _mat3(src)

            return NPKernel(run)

        def zero(self, m):
            def run():
                _mat3(m)[:] = 0

            return NPKernel(run)

        def axnpby(self, *arr, in_scale=(), in_scale_idxs=(), out_scale=()):
            if any(arr[0].traits != x.traits for x in arr[1:]):
                raise ValueError('Incompatible matrix types')
            if in_scale or out_scale:
                raise NotImplementedError('axnpby scaling not in oracle')

            class AxnpbyKernel(NPKernel):
                def bind(self, *consts):
                    self._c = consts

                def run(self, *args):
                    c = self._c
                    xs = [_mat3(x) for x in arr]
                    acc = c[0]*xs[0] if c[0] != 0 else 0
                    for ci, xi in zip(c[1:], xs[1:]):
                        acc = acc + ci*xi
                    xs[0][:] = acc

            return AxnpbyKernel(None)

        def reduction(self, rop, expr, vvars, svars=[], pvars={}):
            """pyfr/backends/base/blasext.py:21-46 evaluated over the
            logical (point, variable, element) arrays: padding never
            enters, per-variable constants broadcast along axis 1."""
            vs = list(vvars.values())
            if any(v.traits != vs[0].traits for v in vs[1:]):
                raise ValueError('Incompatible matrix types')
            if rop not in ('sum', 'max'):
                raise ValueError('Invalid reduction operator')

            fns = {'fabs': np.abs, 'sqrt': np.sqrt, 'fmax': np.maximum,
                   'fmin': np.minimum, 'pow': np.power, 'exp': np.exp}
            pv = {k: np.asarray(v, dtype=vs[0].dtype)[None, :, None]
                  for k, v in pvars.items()}
            red = np.sum if rop == 'sum' else np.max

            class ReductionKernel(NPKernel):
                _sv = ()

                def bind(self, *consts):
                    self._sv = consts

                @property
                def retval(self):
                    return self._ret

                def run(self, *args):
                    ns = dict(fns, **pv, **dict(zip(svars, self._sv)),
                              **{k: v.get() for k, v in vvars.items()})
                    with np.errstate(all='ignore'):
                        self._ret = np.array(
                            [red(eval(e, {'__builtins__': {}}, ns))
                             for e in expr], dtype=float
                        )

            return ReductionKernel(None)

        def pack(self, xv):
            v = _ViewRef(xv.view)
            nr, nc, n = xv.nvrow, xv.nvcol, xv.n
            xm = xv.xchgmat

            def run():
                p = _mat3(xm)[0].reshape(nr*nc, n)
                for r in range(nr):
                    for c in range(nc):
                        p[r*nc + c] = v.load(c, r if nr > 1 else None)

            return NPKernel(run)

        def unpack(self, xm):
            # Received data is consumed in place as an 'mpi' array
            return base.NullKernel()

    class PointwiseProvider:
        def __init__(self, backend):
            self.backend = backend
            self._mods = {}

        def register(self, mod):
            name = mod.rsplit('.', 1)[1]

            if self._mods.setdefault(name, mod) != mod:
                raise RuntimeError(f'Attempt to re-register {name!r} with a '
                                   'different module')

            impl = _pointwise_impls.get(mod)
            if impl is not None and not hasattr(self, name):
                setattr(self, name, lambda *a, **kw: impl(self.backend, *a,
                                                          **kw))

    class NPBackend(base.BaseBackend):
        blocks = False

        const_matrix_cls = NPConstMatrix
        matrix_cls = NPMatrix
        matrix_slice_cls = NPMatrixSlice
        view_cls = NPView
        xchg_matrix_cls = NPXchgMatrix
        xchg_view_cls = NPXchgView
        graph_cls = NPGraph
        ordered_meta_kernel_cls = _meta(ordered_cls)
        unordered_meta_kernel_cls = _meta(unordered_cls)

        def __init__(self, cfg):
            super().__init__(cfg)

            self.alignb = cfg.getint('backend-oracle', 'alignb', 64)
            self.soasz = cfg.getint('backend-oracle', 'soasz', 8)
            self.csubsz = cfg.getint('backend-oracle', 'csubsz', self.soasz)
            self.blocks = cfg.getbool('backend-oracle', 'blocks', False)
            self.extended_mul = cfg.getbool('backend-oracle', 'extended-mul',
                                            False)
            self.mul_log = []

            self.pointwise = PointwiseProvider(self)
            self._providers = [BlasProvider(self), self.pointwise]

        def _malloc_impl(self, nbytes):
            return np.zeros(nbytes, dtype=np.uint8)

        def run_kernels(self, kernels, wait=False):
            for k in kernels:
                k.run()

        def run_graph(self, graph, wait=False):
            graph.run()

        def wait(self):
            pass

    NPBackend.name = name
    NPBackend.kernel_cls = NPKernel
    return NPBackend


# -- pointwise kernel restatements (module path -> builder) -----------------
def _geometry(be, tplargs, npts, smats, rcpdjac, verts, upts, need_rcp):
    """Returns a callable giving (smats[i][j], rcpdjac) as arrays that
    broadcast against (nblocks, npts, nchunks, k)."""
    nd = tplargs['ndims']

    if 'linear' in tplargs['ktype']:
        x = [upts.get()[:, d].reshape(1, npts, 1, 1) for d in range(nd)]

        def geo():
            vv = _stacked(verts, nd)          # (nd, nb, nverts, nch, k)
            V = [[vv[i][:, n:n + 1] for i in range(nd)]
                 for n in range(tplargs['nverts'])]
            s, d = ph.calc_smats_detj(tplargs['jac_exprs'], V, x, nd)
            return s, (1/d if need_rcp else None)
    else:
        def geo():
            sm = _stacked(smats, nd, nlead=nd)   # [i][j] -> (nb,npts,nch,k)
            s = [[sm[i][j] for j in range(nd)] for i in range(nd)]
            return s, (_plain(rcpdjac) if need_rcp else None)

    return geo


def _tflux_euler(be, tplargs, dims, extrns={}, u=None, f=None, smats=None,
                 verts=None, upts=None, **kw):
    nd, nv, c = tplargs['ndims'], tplargs['nvars'], tplargs['c']
    geo = _geometry(be, tplargs, dims[0], smats, None, verts, upts, False)

    def run():
        uu, ff = _stacked(u, nv), _stacked(f, nv, nlead=nd)
        s, _ = geo()
        ft, p, v = ph.inviscid_flux(list(uu), nd, nv, c)
        out = ph.transform_flux(ft, s, nd, nv)
        for i in range(nd):
            for j in range(nv):
                ff[i][j][:] = out[i][j]

    return be.kernel_cls(run)


def _tflux_ns(be, tplargs, dims, extrns={}, u=None, f=None, gradu=None,
              smats=None, rcpdjac=None, verts=None, upts=None,
              artvisc_vtx=None, **kw):
    nd, nv, c = tplargs['ndims'], tplargs['nvars'], tplargs['c']
    fused = 'fused' in tplargs['ktype']
    geo = _geometry(be, tplargs, dims[0], smats, rcpdjac, verts, upts, fused)

    if tplargs.get('shock_capturing', 'none') != 'none':
        raise NotImplementedError('shock capturing is out of scope')

    def run():
        uu, ff = _stacked(u, nv), _stacked(f, nv, nlead=nd)
        s, rcp = geo()

        if fused:
            gg = _stacked(gradu, nv, nlead=nd)
            g = ph.transform_grad([[gg[i][j] for j in range(nv)]
                                   for i in range(nd)], s, rcp, nd, nv)
            for i in range(nd):
                for j in range(nv):
                    gg[i][j][:] = g[i][j]
        else:
            g = [[ff[i][j].copy() for j in range(nv)] for i in range(nd)]

        ft, p, v = ph.inviscid_flux(list(uu), nd, nv, c)
        ph.viscous_flux_add(list(uu), g, ft, nd, nv, c, tplargs['visc_corr'])
        out = ph.transform_flux(ft, s, nd, nv)

        for i in range(nd):
            for j in range(nv):
                ff[i][j][:] = out[i][j]

    return be.kernel_cls(run)


def _gradcoru(be, tplargs, dims, extrns={}, gradu=None, smats=None,
              rcpdjac=None, verts=None, upts=None, **kw):
    nd, nv = tplargs['ndims'], tplargs['nvars']
    geo = _geometry(be, tplargs, dims[0], smats, rcpdjac, verts, upts, True)

    def run():
        gg = _stacked(gradu, nv, nlead=nd)
        s, rcp = geo()
        g = ph.transform_grad([[gg[i][j] for j in range(nv)]
                               for i in range(nd)], s, rcp, nd, nv)
        for i in range(nd):
            for j in range(nv):
                gg[i][j][:] = g[i][j]

    return be.kernel_cls(run)


def _negdivconf(be, tplargs, dims, extrns={}, tdivtconf=None, rcpdjac=None,
                ploc=None, u=None, **kw):
    if tplargs['src_macros']:
        raise NotImplementedError('source terms are out of scope')

    nv = tplargs['nvars']

    def run(t=0.0):
        tt, r = _stacked(tdivtconf, nv), _plain(rcpdjac)
        for i in range(nv):
            tt[i][:] = -r*tt[i]

    return be.kernel_cls(run, rtnames=('t',))


def _wavespeed(be, tplargs, dims, extrns={}, u=None, wspd=None, smats=None,
               rcpdjac=None, verts=None, upts=None, **kw):
    """pyfr/solvers/euler/kernels/wavespeed.mako: per element the maximum
    over its points of sum_i |(S_i/|J|).v| + c |S_i/|J||."""
    nd, nv = tplargs['ndims'], tplargs['nvars']
    gamma = tplargs['c']['gamma']
    geom = _geometry(be, tplargs, dims[0], smats, rcpdjac, verts, upts, True)

    def run(t=0.0):
        sm, rj = geom()
        us = _stacked(u, nv)
        rho = us[0]
        v = [us[i + 1]/rho for i in range(nd)]
        p = (gamma - 1)*(us[nv - 1] - 0.5*rho*sum(vi*vi for vi in v))
        csnd = np.sqrt(gamma*p/rho)

        lam = 0
        for i in range(nd):
            sij = [sm[i][j]*rj for j in range(nd)]
            lam = lam + (np.abs(sum(a*b for a, b in zip(sij, v)))
                         + csnd*np.sqrt(sum(a*a for a in sij)))

        # (nblocks, npts, nchunks, k) -> max over the points
        _plain(wspd)[:] = lam.max(axis=1, keepdims=True)

    return be.kernel_cls(run)


def _rkvdh2(be, tplargs, dims, extrns={}, r1=None, r2=None, rold=None,
            rerr=None, **kw):
    """pyfr/integrators/explicit/kernels/rkvdh2.mako: one stage of a
    two-register van der Houwen Runge-Kutta scheme (Kennedy, Carpenter &
    Lewis 2000), optionally accumulating the embedded error estimate."""
    a, b, e = tplargs['a'], tplargs['b'], tplargs['e']
    stage, nstages = tplargs['stage'], tplargs['nstages']
    errest = tplargs['errest']

    def run(dt=0.0):
        x1, x2 = _mat3(r1), _mat3(r2)
        t1, t2 = x1.copy(), x2.copy()

        if errest and stage == 0:
            _mat3(rerr)[:] = dt*e[stage]*t2
            _mat3(rold)[:] = t1
        elif errest:
            xe = _mat3(rerr)
            xe[:] = xe + dt*e[stage]*t2

        if stage < nstages - 1:
            x1[:] = t1 + dt*a[stage]*t2
            x2[:] = t1 + dt*b[stage]*t2
        else:
            x1[:] = t1 + dt*b[stage]*t2

    return be.kernel_cls(run, rtnames=('dt',))


def _evalsrcmacros(be, tplargs, dims, extrns={}, ploc=None, u=None, **kw):
    """baseadvec/kernels/evalsrcmacros.mako: u <- sum of source macros."""
    if tplargs['src_macros']:
        raise NotImplementedError('source terms are out of scope')

    def run(t=0.0):
        _mat3(u)[:] = 0

    return be.kernel_cls(run, rtnames=('t',))


def _fieldeval(be, tplargs, dims, extrns={}, u=None, gradu=None, ploc=None,
               wts=None, out=None, **kw):
    """pyfr/plugins/kernels/fieldeval.mako (sum / min / max over the
    points of an element, optional weights or mask, coordinates) with
    con_to_pri / grad_con_to_pri of pyfr/solvers/euler/kernels/eos.mako."""
    nd, nv = tplargs['ndims'], tplargs['nvars']
    gm1 = tplargs['c']['gamma'] - 1
    exprs, rop = tplargs['exprs'], tplargs['reduceop']
    has_wts = bool(tplargs.get('has_wts', wts is not None))

    if tplargs.get('use_views') or rop not in ('sum', 'min', 'max'):
        raise NotImplementedError('oracle fieldeval over views')

    fns = {'sqrt': np.sqrt, 'exp': np.exp, 'log': np.log, 'sin': np.sin,
           'cos': np.cos, 'tan': np.tan, 'tanh': np.tanh, 'pow': np.power,
           'fabs': np.abs, 'abs': np.abs, 'fmin': np.minimum,
           'fmax': np.maximum}

    def run(t=0.0):
        cons = list(_stacked(u, nv))
        invrho = 1.0/cons[0]
        rhov = cons[1:nd + 1]
        vel = [invrho*r for r in rhov]
        pri = [cons[0], *vel,
               gm1*(cons[nv - 1] - 0.5*invrho*sum(r*r for r in rhov))]
        env = dict(fns, pri=pri, t=t)

        if tplargs['has_grads']:
            gc = _stacked(gradu, nv, nlead=nd)          # [d][v]
            gp = [[None]*nd for _ in range(nv)]
            for d in range(nd):
                gp[0][d] = gc[d][0]
            for i in range(nd):
                for d in range(nd):
                    gp[i + 1][d] = invrho*(gc[d][i + 1] - vel[i]*gc[d][0])
            for d in range(nd):
                term = 0
                for i in range(nd):
                    term = term + (vel[i]*gc[d][i + 1]
                                   + rhov[i]*gp[i + 1][d])
                gp[nv - 1][d] = gm1*(gc[d][nv - 1] - 0.5*term)
            env['grad_pri'] = gp

        if ploc is not None:
            env['ploc'] = list(_stacked(ploc, nd))

        o = _plain(out)
        w = _plain(wts) if has_wts else None
        fmax = np.finfo(o.dtype).max

        for j, e in enumerate(exprs):
            val = eval(e, {'__builtins__': {}}, env) + 0*cons[0]
            if rop == 'sum':
                o[:, j] = (w*val).sum(axis=1)
            elif rop == 'max':
                val = np.where(w > 0, val, -fmax) if has_wts else val
                o[:, j] = val.max(axis=1)
            else:
                val = np.where(w > 0, val, fmax) if has_wts else val
                o[:, j] = val.min(axis=1)

    return be.kernel_cls(run, rtnames=('t',))


def _mpi_rows(xm, n):
    """A received halo matrix as its ``[nvrow*nvcol][n]`` rows: 'mpi'
    arguments are tightly packed and addressed ``(nv*i + v)*_nx + x``
    (pyfr/backends/base/generator.py:214-226)."""
    a = _mat3(xm)[0]
    assert a.shape[1] == xm.leaddim
    return a.reshape(-1, n)


def _intcflux_euler(be, tplargs, dims, extrns={}, ul=None, ur=None, nl=None,
                    **kw):
    nd, nv, c = tplargs['ndims'], tplargs['nvars'], tplargs['c']
    rs = tplargs['rsolver']
    vl, vr, n = _ViewRef(ul), _ViewRef(ur), list(nl.get())

    def run():
        l = [vl.load(i) for i in range(nv)]
        r = [vr.load(i) for i in range(nv)]
        fn = ph.euler_common_flux(l, r, n, nd, nv, c, rs)
        for i in range(nv):
            vl.store(fn[i], i)
            vr.store(-fn[i], i)

    return be.kernel_cls(run)


def _mpicflux_euler(be, tplargs, dims, extrns={}, ul=None, ur=None, nl=None,
                    **kw):
    nd, nv, c = tplargs['ndims'], tplargs['nvars'], tplargs['c']
    rs = tplargs['rsolver']
    vl, n = _ViewRef(ul), list(nl.get())

    def run():
        l = [vl.load(i) for i in range(nv)]
        r = list(_mpi_rows(ur, dims[0]))
        fn = ph.euler_common_flux(l, r, n, nd, nv, c, rs)
        for i in range(nv):
            vl.store(fn[i], i)

    return be.kernel_cls(run)


def _intconu(be, tplargs, dims, extrns={}, ulin=None, urin=None, ulout=None,
             urout=None, **kw):
    nv, beta = tplargs['nvars'], tplargs['c']['ldg-beta']
    li, ri, lo, ro = map(_ViewRef, (ulin, urin, ulout, urout))

    def run():
        l = [li.load(i) for i in range(nv)]
        r = [ri.load(i) for i in range(nv)]
        ol, orr = ph.ldg_common_solution(l, r, beta)
        for i in range(nv):
            if ol is not None:
                lo.store(ol[i], i)
            if orr is not None:
                ro.store(orr[i], i)

    return be.kernel_cls(run)


def _mpiconu(be, tplargs, dims, extrns={}, ulin=None, urin=None, ulout=None,
             **kw):
    nv, beta = tplargs['nvars'], tplargs['c']['ldg-beta']
    li, lo = _ViewRef(ulin), _ViewRef(ulout)

    def run():
        l = [li.load(i) for i in range(nv)]
        r = list(_mpi_rows(urin, dims[0]))

        # mpiconu.mako:9-18: beta = -1/2 keeps our own trace
        if beta == -0.5:
            out = l
        elif beta == 0.5:
            out = r
        else:
            out = [b*(0.5 + beta) + a*(0.5 - beta) for a, b in zip(l, r)]

        for i in range(nv):
            lo.store(out[i], i)

    return be.kernel_cls(run)


def _cflux_ns(mpi):
    def build(be, tplargs, dims, extrns={}, ul=None, ur=None, gradul=None,
              gradur=None, artvisc=None, nl=None, **kw):
        nd, nv, c = tplargs['ndims'], tplargs['nvars'], tplargs['c']
        rs, vc = tplargs['rsolver'], tplargs['visc_corr']
        beta = c['ldg-beta']
        vl, gl, n = _ViewRef(ul), _ViewRef(gradul), list(nl.get())

        if not mpi:
            vr, gr = _ViewRef(ur), _ViewRef(gradur)

        def run():
            l = [vl.load(i) for i in range(nv)]
            gL = gR = None

            if mpi:
                r = list(_mpi_rows(ur, dims[0]))
                if beta != 0.5:
                    rows = _mpi_rows(gradur, dims[0])
                    gR = [[rows[nv*d + j] for j in range(nv)]
                          for d in range(nd)]
            else:
                r = [vr.load(i) for i in range(nv)]
                if beta != 0.5:
                    gR = [[gr.load(j, d) for j in range(nv)]
                          for d in range(nd)]

            if beta != -0.5:
                gL = [[gl.load(j, d) for j in range(nv)] for d in range(nd)]

            fn = ph.ns_common_flux(l, r, gL, gR, n, nd, nv, c, rs, vc)

            for i in range(nv):
                vl.store(fn[i], i)
                if not mpi:
                    vr.store(-fn[i], i)

        return be.kernel_cls(run)

    return build


def _bc_env(ploc, t):
    env = {'t': t}
    if ploc is not None:
        env['ploc'] = list(ploc.get())
    return env


def _bcconu(be, tplargs, dims, extrns={}, ulin=None, ulout=None, nlin=None,
            ploc=None, **kw):
    """pyfr/solvers/navstokes/kernels/bcconu.mako"""
    nd, nv, c = tplargs['ndims'], tplargs['nvars'], tplargs['c']
    vi, vo, nl = _ViewRef(ulin), _ViewRef(ulout), list(nlin.get())

    def run(t=0.0):
        l = [vi.load(i) for i in range(nv)]
        mag = np.sqrt(sum(x*x for x in nl))
        n = [(1/mag)*x for x in nl]
        ur = ph.bc_ldg_state(tplargs['bctype'], l, n, nd, nv, c,
                             _bc_env(ploc, t))
        for i in range(nv):
            vo.store(ur[i] + 0*l[0], i)

    return be.kernel_cls(run, rtnames=('t',))


def _bccflux(viscous):
    def build(be, tplargs, dims, extrns={}, ul=None, gradul=None,
              artvisc=None, nl=None, ploc=None, **kw):
        """euler/kernels/bccflux.mako, navstokes/kernels/bccflux.mako"""
        nd, nv, c = tplargs['ndims'], tplargs['nvars'], tplargs['c']
        vl, n = _ViewRef(ul), list(nl.get())
        gl = _ViewRef(gradul) if viscous else None

        def run(t=0.0):
            l = [vl.load(i) for i in range(nv)]
            g = ([[gl.load(j, d) for j in range(nv)] for d in range(nd)]
                 if viscous else None)
            fn = ph.bc_common_flux(
                tplargs['bctype'], tplargs.get('bccfluxstate'), l, g, n, nd,
                nv, c, tplargs['rsolver'], _bc_env(ploc, t), viscous,
                tplargs.get('visc_corr', 'none')
            )
            for i in range(nv):
                vl.store(fn[i], i)

        return be.kernel_cls(run, rtnames=('t',))

    return build


_pointwise_impls = {
    'pyfr.solvers.euler.kernels.tflux': _tflux_euler,
    'pyfr.solvers.navstokes.kernels.tflux': _tflux_ns,
    'pyfr.solvers.baseadvecdiff.kernels.gradcoru': _gradcoru,
    'pyfr.solvers.baseadvec.kernels.negdivconf': _negdivconf,
    'pyfr.solvers.baseadvec.kernels.evalsrcmacros': _evalsrcmacros,
    'pyfr.solvers.euler.kernels.intcflux': _intcflux_euler,
    'pyfr.solvers.euler.kernels.mpicflux': _mpicflux_euler,
    'pyfr.solvers.navstokes.kernels.intconu': _intconu,
    'pyfr.solvers.navstokes.kernels.mpiconu': _mpiconu,
    'pyfr.solvers.navstokes.kernels.intcflux': _cflux_ns(False),
    'pyfr.solvers.navstokes.kernels.mpicflux': _cflux_ns(True),
    'pyfr.plugins.kernels.fieldeval': _fieldeval,
    'pyfr.integrators.explicit.kernels.rkvdh2': _rkvdh2,
    'pyfr.solvers.euler.kernels.wavespeed': _wavespeed,
    'pyfr.solvers.navstokes.kernels.bcconu': _bcconu,
    'pyfr.solvers.navstokes.kernels.bccflux': _bccflux(True),
    'pyfr.solvers.euler.kernels.bccflux': _bccflux(False),
}


class LocalComm:
    """All ranks live in one process; a send parks a copy of the buffer in
    a shared mailbox and the matching receive (started at the top of the
    same graph stage on the peer) collects it when the stage is closed by
    ``deliver()``.  Drives multi-partition oracle runs without MPI."""

    def __init__(self, rank, size, world=None):
        self.rank, self.size = rank, size
        self.world = world if world is not None else {'box': {}, 'recv': []}

    def peer(self, rank):
        return LocalComm(rank, self.size, self.world)

    def send_init(self, xm, pid, tag):
        comm = self

        class Send:
            def start(self):
                comm.world['box'][comm.rank, pid, tag] = _mat3(xm).copy()

        return Send()

    def recv_init(self, xm, pid, tag):
        comm = self

        class Recv:
            def start(self):
                comm.world['recv'].append(((pid, comm.rank, tag), xm))

        return Recv()

    def deliver(self):
        for key, xm in self.world['recv']:
            _mat3(xm)[:] = self.world['box'].pop(key)

        self.world['recv'].clear()
