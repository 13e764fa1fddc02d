"""Per-element-type data and kernels of the flux-reconstruction RHS.

Host-side counterpart of the reference's element classes
(``pyfr/solvers/base/elements.py:36-469`` geometry and buffers,
``pyfr/solvers/baseadvec/elements.py:73-127``,
``pyfr/solvers/baseadvecdiff/elements.py:22-116``,
``pyfr/solvers/euler/elements.py:150-206``,
``pyfr/solvers/navstokes/elements.py:33-128``): it owns the mesh geometry
(metric terms, flux-point ordering, face normals) and *declares* the
element kernels by calling ``backend.kernel(...)`` with exactly the names
and keyword arguments the reference uses, so any backend honouring the
``pyfr.backends.base`` contract can be driven by it.
"""

from functools import cached_property

import numpy as np

from pyfr_b200.base import NullKernel


def fuzzy_lexsort(coords, tol=1e-6):
    """Per-element permutation ordering points by x, then y, then z where
    coordinates closer than ``tol`` count as equal.

    ``coords`` is ``(neles, ndims, npts)``; semantics follow the reference's
    ``batched_fuzzysort`` (``pyfr/nputil.py:139-181``) for well separated
    point sets.
    """
    neles, ndims, npts = coords.shape
    perm = np.tile(np.arange(npts), (neles, 1))
    group = np.zeros((neles, npts), dtype=np.int64)

    for d in range(ndims):
        vals = np.take_along_axis(coords[:, d, :], perm, axis=1)

        # Order by (group, value): stable sort on value then on group
        o1 = np.argsort(vals, axis=1, kind='stable')
        g1 = np.take_along_axis(group, o1, axis=1)
        o2 = np.argsort(g1, axis=1, kind='stable')
        o = np.take_along_axis(o1, o2, axis=1)

        perm = np.take_along_axis(perm, o, axis=1)
        vals = np.take_along_axis(vals, o, axis=1)
        group = np.take_along_axis(group, o, axis=1)

        # Split groups where the sorted value jumps by at least tol
        jump = (np.diff(vals, axis=1) >= tol) | (np.diff(group, axis=1) != 0)
        group = np.concatenate(
            [np.zeros((neles, 1), dtype=np.int64), np.cumsum(jump, axis=1)],
            axis=1
        )

    return perm


class BaseElements:
    """Geometry, buffers and kernels for one element type."""

    # Set by the physics subclasses
    system = None

    def __init__(self, shapecls, spts, cfg):
        self.cfg = cfg
        self.eles = spts
        self.nspts, self.neles, self.ndims = spts.shape

        self.basis = basis = shapecls(self.nspts, cfg)
        self.name = basis.name
        self.nupts, self.nfpts = basis.nupts, basis.nfpts
        self.nfacefpts = basis.nfacefpts
        self.nvars = self.ndims + 2
        self.antialias = basis.antialias
        self.nqpts = basis.nqpts if 'flux' in self.antialias else None

        self.kernels = {}
        self._be = None
        self._opmats = {}

    # -- physics ---------------------------------------------------------
    @property
    def privars(self):
        return ['rho', 'u', 'v', 'p'] if self.ndims == 2 else \
               ['rho', 'u', 'v', 'w', 'p']

    def pri_to_con(self, pris):
        rho, p = pris[0], pris[-1]
        gamma = self.cfg.getfloat('constants', 'gamma')

        rhovs = [rho*v for v in pris[1:-1]]
        E = p/(gamma - 1) + 0.5*rho*sum(v*v for v in pris[1:-1])

        return [rho, *rhovs, E]

    def set_ics_from_cfg(self):
        from pyfr_b200.host.exprs import npeval

        vars = self.cfg.items_as('constants', float)
        coords = self.ploc_at_np('upts')
        vars |= dict(zip('xyz', coords.swapaxes(0, 1)))

        ics = [npeval(self.cfg.getexpr('soln-ics', v), vars)
               for v in self.privars]

        out = np.empty((self.nupts, self.nvars, self.neles))
        for i, v in enumerate(self.pri_to_con(ics)):
            out[:, i] = v

        return out

    # -- geometry --------------------------------------------------------
    def _pts(self, name):
        return getattr(self.basis, name) if isinstance(name, str) else name

    def ploc_at_np(self, name):
        """Physical locations ``(npts, ndims, neles)`` of a point set."""
        pts = self._pts(name)
        op = self.basis.sbasis_at(pts)
        x = op @ self.eles.reshape(self.nspts, -1)

        return x.reshape(len(pts), self.neles, self.ndims).swapaxes(1, 2)

    @cached_property
    def _metric_mpts(self):
        """S-matrices ``(ndims, nmpts, ndims, neles)`` and ``|J|``
        ``(nmpts, neles)`` at the metric points (curl-invariant form in 3-D,
        Kopriva, J. Sci. Comput. 26(3) eq. 37; reference
        ``pyfr/solvers/base/elements.py:382-440``)."""
        b, nd, ne = self.basis, self.ndims, self.neles
        mpts, nm = b.mpts, b.nmpts

        D = [b.mbasis_deriv_at(mpts, d) for d in range(nd)]
        xall = self.ploc_at_np('mpts').swapaxes(1, 2)     # (nm, ne, nd)

        smats = np.empty((nd, nm, nd, ne))
        djac = np.empty((nm, ne))

        # Element chunks: the 3-D form holds ~20 temporaries of the size
        # of ``x``; whole-mesh arrays would put the set-up of a 64^3 brick
        # at 20 GB per rank
        step = max(1, (1 << 22) // (nm*nd))
        for e0 in range(0, ne, step):
            sl = slice(e0, min(e0 + step, ne))
            self._metric_chunk(D, np.ascontiguousarray(xall[:, sl]),
                               smats[..., sl], djac[:, sl])

        return smats, djac

    def _metric_chunk(self, D, x, smats, djac):
        nd = self.ndims
        nm, ne = x.shape[:2]

        # dx[d][p, e, i] = d x_i / d xi_d
        dx = [(Dd @ x.reshape(nm, -1)).reshape(nm, ne, nd) for Dd in D]

        if nd == 2:
            a, bb = dx[0][..., 0], dx[0][..., 1]
            c, d = dx[1][..., 0], dx[1][..., 1]

            smats[0, :, 0], smats[0, :, 1] = d, -c
            smats[1, :, 0], smats[1, :, 1] = -bb, a
            djac[:] = a*d - bb*c
        else:
            # T_j = x cross dx/dxi_j, then S_i = (D_j T_k - D_k T_j)/2
            # (component form: np.cross is several times slower on
            # arrays of this size)
            def cross(a, b):
                out = np.empty_like(a)
                out[..., 0] = a[..., 1]*b[..., 2] - a[..., 2]*b[..., 1]
                out[..., 1] = a[..., 2]*b[..., 0] - a[..., 0]*b[..., 2]
                out[..., 2] = a[..., 0]*b[..., 1] - a[..., 1]*b[..., 0]
                return out

            T = [cross(x, dxd) for dxd in dx]
            DT = [[(Dk @ Tj.reshape(nm, -1)).reshape(nm, ne, nd)
                   for Dk in D] for Tj in T]

            for i, (j, k) in enumerate([(1, 2), (2, 0), (0, 1)]):
                s = 0.5*(DT[k][j] - DT[j][k])
                smats[i] = s.swapaxes(1, 2)

            djac[:] = np.einsum('pei,pei->pe', dx[0], cross(dx[1], dx[2]))

        return smats, djac

    def smat_at_np(self, name):
        smats, _ = self._metric_mpts
        m0 = self.basis.mbasis_at(self._pts(name))
        nd = self.ndims

        out = np.array([m0 @ s.reshape(len(m0[0]), -1) for s in smats])
        return out.reshape(nd, -1, nd, self.neles)

    def rcpdjac_at_np(self, name):
        _, djac = self._metric_mpts
        djac = self.basis.mbasis_at(self._pts(name)) @ djac

        if np.any(djac < -1e-5):
            raise RuntimeError('Negative mesh Jacobians detected')

        return 1.0/djac

    @cached_property
    def plocfpts(self):
        return np.ascontiguousarray(self.ploc_at_np('fpts').swapaxes(1, 2))

    @cached_property
    def srtd_face_fpts(self):
        """For every face, the flux-point rows ``(neles, nfacefpts)`` sorted
        by physical location so both sides of an interface agree."""
        out = []

        for ff in self.basis.facefpts:
            ff = np.asarray(ff)
            coords = self.plocfpts[ff].transpose(1, 2, 0)
            out.append(ff[fuzzy_lexsort(coords)])

        return out

    @cached_property
    def _pnorm_fpts(self):
        # |J| J^{-T} n = S^T n at the flux points: (nfpts, neles, ndims)
        smats = self.smat_at_np('fpts')
        pn = np.einsum('lfke,fl->fek', smats, self.basis.norm_fpts)

        if np.any(np.einsum('fek,fek->fe', pn, pn) < 1e-20):
            raise RuntimeError('Zero face normals detected')

        return pn

    # -- backend ---------------------------------------------------------
    def _scratch_bufs(self):
        raise NotImplementedError

    def set_backend(self, be, nonce, linoff):
        self._be = be
        nd, nv, ne = self.ndims, self.nvars, self.neles
        nu, nf = self.nupts, self.nfpts

        self.grad_fusion = not (be.blocks or 'flux' in self.antialias)

        if self.basis.order >= 2:
            self.linoff = -(-linoff // be.csubsz)*be.csubsz
        else:
            self.linoff = ne

        def alloc(ex, shape):
            return be.matrix(shape, extent=nonce + ex, tags={'align'})

        bufs = self._scratch_bufs()

        if 'scal_fpts' in bufs:
            self._scal_fpts = alloc('scal_fpts', (nf, nv, ne))
        if 'scal_qpts' in bufs:
            self._scal_qpts = alloc('scal_qpts', (self.nqpts, nv, ne))
        if 'vect_upts' in bufs:
            self._vect_upts = alloc('vect_upts', (nd, nu, nv, ne))
        if 'vect_qpts' in bufs:
            self._vect_qpts = alloc('vect_qpts', (nd, self.nqpts, nv, ne))
        if 'vect_fpts' in bufs:
            self._vect_fpts = alloc('vect_fpts', (nd, nf, nv, ne))

        if 'comm_fpts' in bufs:
            self._comm_fpts = alloc('comm_fpts', (nf, nv, ne))
        elif 'vect_fpts' in bufs:
            self._comm_fpts = self._vect_fpts.slice(0, nf)

        if 'grad_upts' in bufs and self.grad_fusion:
            self._grad_upts = alloc('grad_upts', (nd, nu, nv, ne))
        elif hasattr(self, '_vect_upts'):
            self._grad_upts = self._vect_upts

        self.scal_upts = []

    def alloc_bank(self, extent, ic=None):
        m = self._be.matrix((self.nupts, self.nvars, self.neles), ic,
                            extent=extent, tags={'align'})
        self.scal_upts.append(m)
        return m

    def opmat(self, expr):
        if expr not in self._opmats:
            self._opmats[expr] = self._be.const_matrix(
                self.basis.opmat(expr), tags={expr, 'align'}
            )

        return self._opmats[expr]

    @property
    def mesh_regions(self):
        off = self.linoff

        if off == 0:
            return {'linear': self.neles}
        elif off >= self.neles:
            return {'curved': self.neles}
        else:
            return {'curved': off, 'linear': self.neles - off}

    def _slice_mat(self, mat, region, ra=None, rb=None):
        if mat is None:
            return None

        off = self.linoff
        if len(mat.ioshape) >= 3:
            off *= mat.ioshape[-2]
        else:
            off = min(off, mat.ncol)

        if region == 'curved':
            return mat.slice(ra, rb, 0, off)
        else:
            return mat.slice(ra, rb, off, mat.ncol)

    def _sliced_kernel(self, kerns):
        kerns = list(kerns)

        if len(kerns) > 1:
            return self._be.unordered_meta_kernel(kerns, [self.linoff])
        else:
            return kerns[0]

    @cached_property
    def upts(self):
        return self._be.const_matrix(self.basis.upts)

    @cached_property
    def qpts(self):
        return self._be.const_matrix(self.basis.qpts)

    def _const(self, key, fn, region=None):
        cache = self.__dict__.setdefault('_constcache', {})

        if key not in cache:
            cache[key] = self._be.const_matrix(fn(), tags={'align'})

        m = cache[key]
        return self._slice_mat(m, region) if region else m

    def rcpdjac_at(self, name, region=None):
        return self._const(('rcpdjac', name),
                           lambda: self.rcpdjac_at_np(name), region)

    def ploc_at(self, name, region=None):
        return self._const(('ploc', name), lambda: self.ploc_at_np(name),
                           region)

    def curved_smat_at(self, name):
        return self._const(
            ('smat', name),
            lambda: self.smat_at_np(name)[..., :self.linoff]
        )

    # -- interface hooks (what views over our buffers look like) ----------
    def get_pnorms_for_inters(self, eidxs, fidx):
        rows = self.srtd_face_fpts[fidx][eidxs]
        return self._pnorm_fpts[rows, eidxs[:, None]].reshape(-1, self.ndims)

    def get_ploc_for_inters(self, eidxs, fidx):
        rows = self.srtd_face_fpts[fidx][eidxs]
        return self.plocfpts[rows, eidxs[:, None]].reshape(-1, self.ndims)

    def get_scal_fpts_for_inters(self, eidxs, fidx):
        return self._scal_fpts.mid, self.srtd_face_fpts[fidx][eidxs], None

    def get_comm_fpts_for_inters(self, eidxs, fidx):
        if self.basis.fpts_in_upts:
            return (self._comm_fpts.mid, self.srtd_face_fpts[fidx][eidxs],
                    None)
        else:
            return self.get_vect_fpts_for_inters(eidxs, fidx)

    def get_vect_fpts_for_inters(self, eidxs, fidx):
        rows = self.srtd_face_fpts[fidx][eidxs]

        if self.basis.fpts_in_upts and self.grad_fusion:
            return (self._grad_upts.mid, self.basis.fpts_map_upts[rows],
                    self.nupts)
        elif self.basis.fpts_in_upts and not hasattr(self, '_vect_fpts'):
            return (self._vect_upts.mid, self.basis.fpts_map_upts[rows],
                    self.nupts)
        else:
            return self._vect_fpts.mid, rows, self.nfpts


class AdvectionElements(BaseElements):
    def _scratch_bufs(self):
        if 'flux' in self.antialias:
            return {'scal_fpts', 'scal_qpts', 'vect_qpts'}
        else:
            return {'scal_fpts', 'vect_upts'}

    def set_backend(self, be, nonce, linoff):
        super().set_backend(be, nonce, linoff)
        k, order = self.kernels, self.basis.order

        be.pointwise.register('pyfr.solvers.baseadvec.kernels.negdivconf')

        k['disu'] = lambda uin: be.kernel(
            'mul', self.opmat('M0'), self.scal_upts[uin], out=self._scal_fpts
        )

        # Flux anti-aliasing: flux evaluated at the quadrature points and
        # projected back (pyfr/solvers/baseadvec/elements.py:78-94)
        fluxaa = 'flux' in self.antialias

        if fluxaa and order > 0:
            k['qptsu'] = lambda uin: be.kernel(
                'mul', self.opmat('M7'), self.scal_upts[uin],
                out=self._scal_qpts
            )
            k['tdivtpcorf'] = lambda fout: be.kernel(
                'mul', self.opmat('(M1 - M3*M2)*M9'), self._vect_qpts,
                out=self.scal_upts[fout]
            )
        elif order > 0:
            k['tdivtpcorf'] = lambda fout: be.kernel(
                'mul', self.opmat('M1 - M3*M2'), self._vect_upts,
                out=self.scal_upts[fout]
            )

        k['tdivtconf'] = lambda fout: be.kernel(
            'mul', self.opmat('M3'), self._scal_fpts,
            out=self.scal_upts[fout], beta=float(order > 0)
        )

        srctplargs = {'ndims': self.ndims, 'nvars': self.nvars,
                      'src_macros': []}
        k['negdivconf'] = lambda fout: be.kernel(
            'negdivconf', tplargs=srctplargs, dims=[self.nupts, self.neles],
            extrns={}, tdivtconf=self.scal_upts[fout],
            rcpdjac=self.rcpdjac_at('upts'), ploc=None, u=None
        )

    def init_wavespeed(self):
        """Per-element maximum wave speed for the CFL controller
        (pyfr/solvers/euler/elements.py:24-69).  Unlike the reference the
        output is sliced per region, so that on a mesh with curved *and*
        linear elements each kernel writes its own columns."""
        be = self._be
        be.pointwise.register('pyfr.solvers.euler.kernels.wavespeed')

        self._wspd = be.matrix((1, self.neles), tags={'align'})
        tplargs = self._flux_tplargs()
        r, s = self.mesh_regions, self._slice_mat

        def kern(uin):
            ks = []
            if 'curved' in r:
                ks.append(be.kernel(
                    'wavespeed', tplargs=tplargs | {'ktype': 'curved'},
                    dims=[self.nupts, r['curved']],
                    u=s(self.scal_upts[uin], 'curved'),
                    wspd=s(self._wspd, 'curved'),
                    smats=self.curved_smat_at('upts'),
                    rcpdjac=self.rcpdjac_at('upts', 'curved')
                ))
            if 'linear' in r:
                ks.append(be.kernel(
                    'wavespeed', tplargs=tplargs | {'ktype': 'linear'},
                    dims=[self.nupts, r['linear']],
                    u=s(self.scal_upts[uin], 'linear'),
                    wspd=s(self._wspd, 'linear'),
                    verts=self.ploc_at('linspts', 'linear'), upts=self.upts
                ))
            return self._sliced_kernel(ks)

        self.kernels['wavespeed'] = lambda uin: kern(uin)
        return self._wspd

    def _flux_tplargs(self):
        return {
            'ndims': self.ndims, 'nvars': self.nvars,
            'nverts': len(self.basis.linspts),
            'c': self.cfg.items_as('constants', float),
            'jac_exprs': self.basis.jac_exprs
        }


class EulerElements(AdvectionElements):
    system = 'euler'

    def set_backend(self, be, nonce, linoff):
        super().set_backend(be, nonce, linoff)

        if self.basis.order == 0:
            return

        be.pointwise.register('pyfr.solvers.euler.kernels.tflux')
        tplargs = self._flux_tplargs()
        r, s = self.mesh_regions, self._slice_mat

        tdisf = []
        if 'flux' in self.antialias:
            # pyfr/solvers/euler/elements.py:180-206
            if 'curved' in r:
                tdisf.append(lambda: be.kernel(
                    'tflux', tplargs=tplargs | {'ktype': 'curved'},
                    dims=[self.nqpts, r['curved']],
                    u=s(self._scal_qpts, 'curved'),
                    f=s(self._vect_qpts, 'curved'),
                    smats=self.curved_smat_at('qpts')
                ))
            if 'linear' in r:
                tdisf.append(lambda: be.kernel(
                    'tflux', tplargs=tplargs | {'ktype': 'linear'},
                    dims=[self.nqpts, r['linear']],
                    u=s(self._scal_qpts, 'linear'),
                    f=s(self._vect_qpts, 'linear'),
                    verts=self.ploc_at('linspts', 'linear'), upts=self.qpts
                ))

            self.kernels['tdisf'] = lambda: self._sliced_kernel(
                k() for k in tdisf
            )
            return

        if 'curved' in r:
            tdisf.append(lambda uin: be.kernel(
                'tflux', tplargs=tplargs | {'ktype': 'curved'},
                dims=[self.nupts, r['curved']],
                u=s(self.scal_upts[uin], 'curved'),
                f=s(self._vect_upts, 'curved'),
                smats=self.curved_smat_at('upts')
            ))
        if 'linear' in r:
            tdisf.append(lambda uin: be.kernel(
                'tflux', tplargs=tplargs | {'ktype': 'linear'},
                dims=[self.nupts, r['linear']],
                u=s(self.scal_upts[uin], 'linear'),
                f=s(self._vect_upts, 'linear'),
                verts=self.ploc_at('linspts', 'linear'), upts=self.upts
            ))

        self.kernels['tdisf'] = lambda uin: self._sliced_kernel(
            k(uin) for k in tdisf
        )


class NavierStokesElements(AdvectionElements):
    system = 'navier-stokes'

    def _scratch_bufs(self):
        bufs = {'scal_fpts', 'vect_fpts', 'vect_upts'}

        if 'flux' in self.antialias:
            bufs |= {'scal_qpts', 'vect_qpts'}
        elif self.grad_fusion:
            bufs |= {'grad_upts'}

        if self.basis.fpts_in_upts:
            bufs |= {'comm_fpts'}
            if self.grad_fusion:
                bufs -= {'vect_fpts'}

        return bufs

    def set_backend(self, be, nonce, linoff):
        super().set_backend(be, nonce, linoff)

        kernel, k = be.kernel, self.kernels
        order, nu, nf = self.basis.order, self.nupts, self.nfpts
        r, s = self.mesh_regions, self._slice_mat

        be.pointwise.register('pyfr.solvers.baseadvecdiff.kernels.gradcoru')

        if abs(self.cfg.getfloat('solver-interfaces', 'ldg-beta')) == 0.5:
            k['copy_fpts'] = lambda: kernel('copy', self._comm_fpts,
                                            self._scal_fpts)

        if order > 0:
            k['tgradpcoru_upts'] = lambda uin: kernel(
                'mul', self.opmat('M4 - M6*M0'), self.scal_upts[uin],
                out=self._grad_upts
            )

        k['tgradcoru_upts'] = lambda: kernel(
            'mul', self.opmat('M6'), self._comm_fpts, out=self._grad_upts,
            beta=float(order > 0)
        )

        gtpl = {'ndims': self.ndims, 'nvars': self.nvars,
                'nverts': len(self.basis.linspts),
                'jac_exprs': self.basis.jac_exprs}

        gradcoru_u = []
        if 'curved' in r:
            gradcoru_u.append(lambda: kernel(
                'gradcoru', tplargs=gtpl | {'ktype': 'curved'},
                dims=[nu, r['curved']], gradu=s(self._grad_upts, 'curved'),
                smats=self.curved_smat_at('upts'),
                rcpdjac=self.rcpdjac_at('upts', 'curved')
            ))
        if 'linear' in r:
            gradcoru_u.append(lambda: kernel(
                'gradcoru', tplargs=gtpl | {'ktype': 'linear'},
                dims=[nu, r['linear']], gradu=s(self._grad_upts, 'linear'),
                upts=self.upts, verts=self.ploc_at('linspts', 'linear')
            ))

        k['gradcoru_u'] = lambda: self._sliced_kernel(g() for g in gradcoru_u)

        if not self.grad_fusion or order == 0:
            k['gradcoru_upts'] = k['gradcoru_u']

        def gradcoru_fpts():
            vu, vf = self._grad_upts, self._vect_fpts
            muls = [kernel('mul', self.opmat('M0'),
                           vu.slice(i*nu, (i + 1)*nu),
                           vf.slice(i*nf, (i + 1)*nf))
                    for i in range(self.ndims)]

            return be.unordered_meta_kernel(muls)

        if not (self.basis.fpts_in_upts and self.grad_fusion):
            k['gradcoru_fpts'] = gradcoru_fpts

        fluxaa = 'flux' in self.antialias

        if fluxaa and order > 0:
            # pyfr/solvers/baseadvecdiff/elements.py:103-116
            def gradcoru_qpts():
                nq = self.nqpts
                vu, vq = self._vect_upts, self._vect_qpts
                muls = [kernel('mul', self.opmat('M7'),
                               vu.slice(i*nu, (i + 1)*nu),
                               vq.slice(i*nq, (i + 1)*nq))
                        for i in range(self.ndims)]

                return be.unordered_meta_kernel(muls)

            k['gradcoru_qpts'] = gradcoru_qpts

        if order == 0:
            return

        be.pointwise.register('pyfr.solvers.navstokes.kernels.tflux')

        visc_corr = self.cfg.get('solver', 'viscosity-correction', 'none')
        if visc_corr not in {'sutherland', 'none'}:
            raise ValueError('Invalid viscosity-correction option')

        tplargs = self._flux_tplargs() | {
            'shock_capturing': self.cfg.get('solver', 'shock-capturing',
                                            'none'),
            'visc_corr': visc_corr
        }

        fused = self.grad_fusion
        kname = 'tdisf_fused' if fused else 'tdisf'

        if fluxaa:
            # Flux at the quadrature points (navstokes/elements.py:64-128)
            specs = []
            for rgn in ('curved', 'linear'):
                if rgn not in r:
                    continue

                kw = ({'smats': self.curved_smat_at('qpts')}
                      if rgn == 'curved' else
                      {'verts': self.ploc_at('linspts', 'linear')})
                kw |= dict(upts=self.qpts, u=s(self._scal_qpts, rgn),
                           f=s(self._vect_qpts, rgn))
                specs.append((rgn, r[rgn], kw))

            k['tdisf'] = lambda: self._sliced_kernel(
                kernel('tflux', tplargs=tplargs | {'ktype': kt},
                       dims=[self.nqpts, n], artvisc_vtx=None, **kw)
                for kt, n, kw in specs
            )
            return

        specs = []
        for rgn in ('curved', 'linear'):
            if rgn not in r:
                continue

            kw = {}
            if rgn == 'curved':
                kw['smats'] = self.curved_smat_at('upts')
                if fused:
                    kw['rcpdjac'] = self.rcpdjac_at('upts', 'curved')
            else:
                kw['verts'] = self.ploc_at('linspts', 'linear')

            kw['upts'] = self.upts
            kw['f'] = s(self._vect_upts, rgn)
            if fused:
                kw['gradu'] = s(self._grad_upts, rgn)

            specs.append((f'{rgn}-fused' if fused else rgn, r[rgn], rgn, kw))

        k[kname] = lambda uin: self._sliced_kernel(
            kernel('tflux', tplargs=tplargs | {'ktype': kt}, dims=[nu, n],
                   u=s(self.scal_upts[uin], rgn), artvisc_vtx=None, **kw)
            for kt, n, rgn, kw in specs
        )
