"""RHS orchestration: kernel collection and the 2/3 graphs per evaluation.

Host-side counterpart of ``pyfr/solvers/base/system.py:15-494``
(``BaseSystem``: element/interface loading, register banks, kernel
instantiation per bank, ``rhs``), ``pyfr/solvers/baseadvec/system.py:58-136``
(Euler: 2 graphs) and ``pyfr/solvers/baseadvecdiff/system.py:54-231``
(Navier-Stokes: 3 graphs, with the block-fusion groups handed to
``Graph.group``).  Everything numerical happens in backend kernels reached
through the ``pyfr.backends.base`` call surface.
"""

from collections import defaultdict
import inspect
import itertools as it

import numpy as np

from pyfr_b200.base import NullKernel
from pyfr_b200.host.elements import EulerElements, NavierStokesElements
from pyfr_b200.host.inters import (EulerIntInters, EulerMPIInters,
                                   NavierStokesIntInters,
                                   NavierStokesMPIInters, euler_bc_map,
                                   navstokes_bc_map)
from pyfr_b200.host.shapes import shape_map


class SerialComm:
    rank, size = 0, 1


class BaseSystem:
    name = None
    elementscls = intinterscls = mpiinterscls = None
    bcmap = {}
    _nonces = it.count()

    def __init__(self, backend, mesh, initsoln, nregs, cfg, comm=None,
                 needs_cfl=False):
        self.backend = be = backend
        self._needs_cfl = needs_cfl
        self.mesh, self.cfg = mesh, cfg
        self.comm = comm = comm or SerialComm()
        self.ndims = mesh.ndims
        self.nonce = nonce = str(next(self._nonces))

        # Elements
        self.ele_map = elemap = {
            et: self.elementscls(shape_map[et], spts, cfg)
            for et, spts in mesh.spts.items()
        }
        eles = list(elemap.values())
        self.nvars = eles[0].nvars

        if initsoln is not None:
            ics = [initsoln[et] for et in elemap]
        else:
            ics = [e.set_ics_from_cfg() for e in eles]

        for et, e in elemap.items():
            curved = mesh.spts_curved[et]
            linoff = np.max(np.nonzero(curved)[0], initial=-1) + 1
            e.set_backend(be, nonce, linoff)

        # pyfr/solvers/baseadvec/system.py:16-18
        if needs_cfl:
            for e in eles:
                e.init_wavespeed()

        be.commit()

        self.ele_types = list(elemap)
        self.ele_ndofs = [e.neles*e.nupts*e.nvars for e in eles]
        self.ele_shapes = {et: (e.nupts, e.nvars, e.neles)
                           for et, e in elemap.items()}

        # Register banks (all hold the initial condition, as in the
        # reference where RHS banks are seeded with ics)
        self.nrhs = nregs
        self.ele_banks = [[e.alloc_bank(f'bank_{nonce}_{i}', ic=ic)
                           for _ in range(nregs)]
                          for i, (e, ic) in enumerate(zip(eles, ics))]

        # Interfaces
        self._int_inters = [self.intinterscls(be, *mesh.con, elemap, cfg)]
        self._mpi_inters = [
            self.mpiinterscls(be, con, p, comm.rank, elemap, cfg)
            for p, con in mesh.con_p.items()
        ]

        # Boundaries, in codec order (pyfr/solvers/base/system.py:262-306)
        self._bc_inters = []
        for c in mesh.codec:
            if c.startswith('bc/') and c[3:] in mesh.bcon:
                sect = f'soln-bcs-{c[3:]}'
                cls = self.bcmap.get(cfg.get(sect, 'type'))
                if cls is None:
                    raise NotImplementedError(
                        f'Boundary type {cfg.get(sect, "type")!r} is not on '
                        'this path (see DESIGN.md)'
                    )
                self._bc_inters.append(cls(be, mesh.bcon[c[3:]], elemap,
                                           sect, cfg))

    # -- exchange registration -------------------------------------------
    def register_mpi_exchange(self, name, views, send=None, recv=None):
        be, comm = self.backend, self.comm

        def reg(m, lhs, rhs, tag):
            if not send or send(m):
                m.kernels[f'{name}_pack'] = lambda: be.kernel('pack', lhs)
                m.mpireqs[f'{name}_send'] = lambda: lhs.sendreq(
                    comm, m.rhsrank, tag
                )
            if not recv or recv(m):
                m.kernels[f'{name}_unpack'] = lambda: be.kernel('unpack', rhs)
                m.mpireqs[f'{name}_recv'] = lambda: rhs.recvreq(
                    comm, m.rhsrank, tag
                )

        for m, (lhs, rhs) in zip(self._mpi_inters, views):
            reg(m, lhs, rhs, m.next_mpi_tag())

    def commit(self):
        self.register_mpi_exchange(
            'scal_fpts', [(m.scal_lhs, m.scal_rhs) for m in self._mpi_inters]
        )
        self.backend.commit()

        self._gen_kernels()

        # Reduction kernels for the largest wave speed of each element type
        # (pyfr/solvers/baseadvec/system.py:48-54)
        if self._needs_cfl:
            self._wspd_red_kerns = [
                self.backend.kernel('reduction', 'max', ['x'], {'x': e._wspd})
                for e in self.ele_map.values()
            ]

        self.backend.commit()

        # What the post-step field integrator needs of the elements
        # (reference: intg.system.ele_map / eles_vect_upts)
        self.eles_vect_upts = [getattr(e, '_grad_upts', None)
                               for e in self.ele_map.values()]
        self.ele_quad = [(e.basis.upts_wts, e.rcpdjac_at_np('upts'),
                          e.privars, e.ndims, e.nvars, e.nupts, e.neles)
                         for e in self.ele_map.values()]

        # Physical solution-point locations on demand (coordinate-dependent
        # integrands); only the vertices and the small interpolation
        # operator outlive the element objects
        def ploc_getter(e):
            op, x = e.basis.sbasis_at(e.basis.upts), e.eles
            shp = (len(op), e.neles, e.ndims)
            return lambda: (op @ x.reshape(len(x), -1)).reshape(
                shp).swapaxes(1, 2)

        self.ele_ploc_upts = [ploc_getter(e) for e in self.ele_map.values()]

        del self.ele_map, self._int_inters, self._mpi_inters
        del self._bc_inters
        self._graphs = {}
        self._ggraphs = {}

    def _gen_kernels(self):
        self._kernels = kernels = defaultdict(list)
        self._ktags = {}
        self._mpireqs = mpireqs = defaultdict(list)

        groups = [('eles', self.ele_map.values()), ('iint', self._int_inters),
                  ('mpiint', self._mpi_inters), ('bcint', self._bc_inters)]

        for pn, provs in groups:
            for p in provs:
                for kn, getter in p.kernels.items():
                    params = inspect.signature(getter).parameters

                    if 'uin' in params or 'fout' in params:
                        for i in range(self.nrhs):
                            kern = getter(i)
                            if isinstance(kern, NullKernel):
                                continue

                            key = ((f'{pn}/{kn}', i, None) if 'uin' in params
                                   else (f'{pn}/{kn}', None, i))
                            kernels[key].append(kern)
                            self._ktags[kern] = f'{pn}/{p.name}'
                    else:
                        kern = getter()
                        if isinstance(kern, NullKernel):
                            continue

                        kernels[f'{pn}/{kn}', None, None].append(kern)
                        self._ktags[kern] = f'{pn}/{p.name}'

        for m in self._mpi_inters:
            for mn, getter in m.mpireqs.items():
                mpireqs[mn].append(getter())

    def _get_kernels(self, uin, fout):
        out = defaultdict(list)

        for (kn, ui, fo), ks in self._kernels.items():
            if ((ui is None and fo is None) or
                (ui is not None and ui == uin) or
                (fo is not None and fo == fout)):
                out[kn].extend(ks)

        return out

    def _kdeps(self, kdict, kern, *names):
        tag = self._ktags[kern]
        return [k for n in names for k in kdict[n] if self._ktags[k] == tag]

    def _group(self, g, kerns, subs=[]):
        kerns = [k for k in kerns if k is not None]
        subs = [[(k, n) for k, n in sub if k] for sub in subs]
        g.group(kerns, [sub for sub in subs if len(sub) > 1])

    def _rhs_graphs(self, uin, fout, post=()):
        raise NotImplementedError

    def _add_post(self, g, k, post):
        """Appends per-element-type kernels that consume the finished RHS
        (a Runge-Kutta stage update) to its last graph, grouped with the
        kernels they follow so that the backend may fuse them."""
        for l in post:
            g.add(l, deps=k['eles/negdivconf'])

    def rhs_graphs(self, uin, fout, post=None):
        """``post = (key, kernels)``: see ``_add_post``; graphs are
        memoised per ``(uin, fout, key)``."""
        key = (uin, fout) if post is None else (uin, fout, post[0])

        if key not in self._graphs:
            self._graphs[key] = (
                self._rhs_graphs(uin, fout) if post is None else
                self._rhs_graphs(uin, fout, tuple(post[1]))
            )

        return self._graphs[key]

    def rhs(self, t, uinbank, foutbank, post=None, **rtargs):
        """``rtargs``: run-time scalars of the ``post`` kernels (``dt``)."""
        if uinbank >= self.nrhs or foutbank >= self.nrhs:
            raise ValueError('Invalid register numbers')

        graphs = self.rhs_graphs(uinbank, foutbank, post)

        for ks in self._get_kernels(uinbank, foutbank).values():
            for k in ks:
                if k.rtnames:
                    k.bind(t=t)

        if post is not None:
            # After fusion the scalars belong to whatever kernel absorbed
            # the post kernels: look them up in the committed plans
            for g in graphs:
                plan = getattr(g, 'plan', None)
                ks = ([k for w, k in plan if w == 'kernel']
                      if plan is not None else post[1])
                for k in ks:
                    if set(getattr(k, 'rtnames', ()) or ()) & set(rtargs):
                        k.bind(**rtargs)

        for g in graphs:
            self.backend.run_graph(g)

    def compute_max_wavespeed(self, uinbank):
        """pyfr/solvers/baseadvec/system.py:157-161"""
        k = self._get_kernels(uinbank, None)
        kerns = k['eles/wavespeed'] + self._wspd_red_kerns
        self.backend.run_kernels(kerns, wait=True)
        return max(float(k.retval[0]) for k in self._wspd_red_kerns)

    def ele_scal_upts(self, idx):
        return [eb[idx].get() for eb in self.ele_banks]


class EulerSystem(BaseSystem):
    name = 'euler'
    elementscls = EulerElements
    intinterscls = EulerIntInters
    mpiinterscls = EulerMPIInters
    bcmap = euler_bc_map

    def _rhs_graphs(self, uin, fout, post=()):
        m, k = self._mpireqs, self._get_kernels(uin, fout)
        deps = lambda dk, *names: self._kdeps(k, dk, *names)
        be = self.backend

        g1 = be.graph()
        g1.add_mpi_reqs(m['scal_fpts_recv'])
        g1.add_all(k['eles/disu'])
        g1.add_all(k['mpiint/scal_fpts_pack'], deps=k['eles/disu'])
        for send, pack in zip(m['scal_fpts_send'],
                              k['mpiint/scal_fpts_pack']):
            g1.add_mpi_req(send, deps=[pack])
        g1.add_all(k['iint/comm_flux'],
                   deps=k['eles/disu'] + k['mpiint/scal_fpts_pack'])
        g1.add_all(k['bcint/comm_flux'], deps=k['eles/disu'])
        g1.commit()

        g2 = be.graph()
        g2.add_all(k['eles/qptsu'])
        for l in k['eles/tdisf']:
            g2.add(l, deps=deps(l, 'eles/qptsu'))
        for l in k['eles/tdivtpcorf']:
            g2.add(l, deps=deps(l, 'eles/tdisf'))
        g2.add_all(k['mpiint/scal_fpts_unpack'])
        for l in k['mpiint/comm_flux']:
            g2.add(l, deps=deps(l, 'mpiint/scal_fpts_unpack'))
        for l in k['eles/tdivtconf']:
            g2.add(l, deps=deps(l, 'eles/tdivtpcorf') + k['mpiint/comm_flux'])
        for l in k['eles/negdivconf']:
            g2.add(l, deps=deps(l, 'eles/tdivtconf'))

        self._add_post(g2, k, post)

        kgroup = [k['eles/qptsu'], k['eles/tdisf'], k['eles/tdivtpcorf'],
                  k['eles/tdivtconf'], k['eles/negdivconf'], list(post)]
        for ks in it.zip_longest(*kgroup):
            self._group(g2, ks, subs=[[(ks[0], 'out'), (ks[1], 'u')],
                                      [(ks[1], 'f'), (ks[2], 'b')]])

        g2.commit()

        return g1, g2


class NavierStokesSystem(BaseSystem):
    name = 'navier-stokes'
    elementscls = NavierStokesElements
    intinterscls = NavierStokesIntInters
    mpiinterscls = NavierStokesMPIInters
    bcmap = navstokes_bc_map

    def commit(self):
        self.register_mpi_exchange(
            'vect_fpts', [(m._vect_lhs, m._vect_rhs)
                          for m in self._mpi_inters],
            send=lambda m: m.c['ldg-beta'] != -0.5,
            recv=lambda m: m.c['ldg-beta'] != 0.5
        )

        # As in the reference the gradient exchange is registered before the
        # solution exchange, so it takes the lower tag on every interface
        super().commit()

    def _rhs_graphs(self, uin, fout, post=()):
        m, k = self._mpireqs, self._get_kernels(uin, fout)
        deps = lambda dk, *names: self._kdeps(k, dk, *names)
        be = self.backend

        # Interpolate to the flux points, exchange, common solution
        g1 = be.graph()
        g1.add_mpi_reqs(m['scal_fpts_recv'])
        g1.add_all(k['eles/disu'])
        g1.add_all(k['mpiint/scal_fpts_pack'], deps=k['eles/disu'])
        for send, pack in zip(m['scal_fpts_send'],
                              k['mpiint/scal_fpts_pack']):
            g1.add_mpi_req(send, deps=[pack])
        for l in k['eles/copy_fpts']:
            g1.add(l, deps=deps(l, 'eles/disu'))
        kd = k['eles/copy_fpts'] or k['eles/disu']
        g1.add_all(k['iint/con_u'], deps=kd + k['mpiint/scal_fpts_pack'])
        g1.add_all(k['bcint/con_u'], deps=kd)
        g1.commit()

        # Gradients, flux, partial divergence
        g2 = be.graph()
        g2.add_mpi_reqs(m['vect_fpts_recv'])
        g2.add_all(k['mpiint/scal_fpts_unpack'])
        for l in k['mpiint/con_u']:
            g2.add(l, deps=deps(l, 'mpiint/scal_fpts_unpack'))
        g2.add_all(k['eles/tgradpcoru_upts'])
        for l in k['eles/tgradcoru_upts']:
            g2.add(l, deps=deps(l, 'eles/tgradpcoru_upts') + k['mpiint/con_u'])
        for l in k['eles/gradcoru_upts']:
            g2.add(l, deps=deps(l, 'eles/tgradcoru_upts'))
        for l in k['eles/tdisf_fused']:
            g2.add(l, deps=deps(l, 'eles/tgradcoru_upts'))
        for l in k['eles/gradcoru_fpts']:
            g2.add(l, deps=deps(l, 'eles/tdisf_fused', 'eles/gradcoru_upts'))

        ideps = k['eles/gradcoru_fpts'] or k['eles/tdisf_fused']

        g2.add_all(k['mpiint/vect_fpts_pack'], deps=ideps)
        for send, pack in zip(m['vect_fpts_send'],
                              k['mpiint/vect_fpts_pack']):
            g2.add_mpi_req(send, deps=[pack])

        g2.add_all(k['iint/comm_flux'], deps=ideps,
                   pdeps=k['mpiint/vect_fpts_pack'])
        g2.add_all(k['bcint/comm_flux'], deps=ideps,
                   pdeps=k['mpiint/vect_fpts_pack'])

        for l in k['eles/gradcoru_qpts']:
            g2.add(l, deps=deps(l, 'eles/gradcoru_upts'),
                   pdeps=k['mpiint/vect_fpts_pack'])

        g2.add_all(k['eles/qptsu'])

        for l in k['eles/tdisf']:
            if k['eles/qptsu']:
                ld = deps(l, 'eles/gradcoru_qpts', 'eles/qptsu')
            elif k['eles/gradcoru_fpts']:
                ld = deps(l, 'eles/gradcoru_fpts')
            else:
                ld = deps(l, 'eles/gradcoru_upts')
            g2.add(l, deps=ld)

        for l in k['eles/tdivtpcorf']:
            g2.add(l, deps=deps(l, 'eles/tdisf', 'eles/tdisf_fused'))

        kgroup = [k['eles/tgradpcoru_upts'], k['eles/tgradcoru_upts'],
                  k['eles/gradcoru_upts'], k['eles/tdisf_fused'],
                  k['eles/gradcoru_fpts'], k['eles/gradcoru_qpts'],
                  k['eles/qptsu'], k['eles/tdisf'], k['eles/tdivtpcorf']]
        for ks in it.zip_longest(*kgroup):
            if k['eles/qptsu']:
                subs = [[(ks[0], 'out'), (ks[1], 'out'), (ks[2], 'gradu'),
                         (ks[4], 'b'), (ks[5], 'b')],
                        [(ks[6], 'out'), (ks[7], 'u')],
                        [(ks[5], 'out'), (ks[7], 'f'), (ks[8], 'b')]]
            elif k['eles/tdisf_fused']:
                subs = [[(ks[0], 'out'), (ks[1], 'out'), (ks[3], 'gradu'),
                         (ks[4], 'b')],
                        [(ks[3], 'f'), (ks[8], 'b')]]
            else:
                subs = [[(ks[0], 'out'), (ks[1], 'out'), (ks[2], 'gradu'),
                         (ks[4], 'b'), (ks[7], 'f'), (ks[8], 'b')]]

            self._group(g2, ks, subs=subs)

        g2.commit()

        # Inter-partition flux, final correction, physical divergence
        g3 = be.graph()
        g3.add_all(k['mpiint/vect_fpts_unpack'])
        for l in k['mpiint/comm_flux']:
            g3.add(l, deps=deps(l, 'mpiint/vect_fpts_unpack'))
        g3.add_all(k['eles/tdivtconf'], deps=k['mpiint/comm_flux'])
        for l in k['eles/negdivconf']:
            g3.add(l, deps=deps(l, 'eles/tdivtconf'))
        self._add_post(g3, k, post)
        for ks in it.zip_longest(k['eles/tdivtconf'], k['eles/negdivconf'],
                                 list(post)):
            self._group(g3, list(ks))
        g3.commit()

        return g1, g2, g3


    def _compute_grads_graph(self, uin):
        """Physical gradients of bank ``uin`` at the solution points, left
        in ``eles_vect_upts`` (pyfr/solvers/baseadvecdiff/system.py:233-286;
        used by post-step plugins such as ``integrate``)."""
        m, k = self._mpireqs, self._get_kernels(uin, None)
        deps = lambda dk, *names: self._kdeps(k, dk, *names)
        be = self.backend

        g1 = be.graph()
        g1.add_mpi_reqs(m['scal_fpts_recv'])
        g1.add_all(k['eles/disu'])
        g1.add_all(k['mpiint/scal_fpts_pack'], deps=k['eles/disu'])
        for send, pack in zip(m['scal_fpts_send'],
                              k['mpiint/scal_fpts_pack']):
            g1.add_mpi_req(send, deps=[pack])
        for l in k['eles/copy_fpts']:
            g1.add(l, deps=deps(l, 'eles/disu'))
        kd = k['eles/copy_fpts'] or k['eles/disu']
        g1.add_all(k['iint/con_u'], deps=kd)
        g1.add_all(k['bcint/con_u'], deps=kd)
        g1.add_all(k['eles/tgradpcoru_upts'],
                   deps=k['iint/con_u'] + k['bcint/con_u'])
        g1.commit()

        g2 = be.graph()
        g2.add_all(k['mpiint/scal_fpts_unpack'])
        for l in k['mpiint/con_u']:
            g2.add(l, deps=deps(l, 'mpiint/scal_fpts_unpack'))
        g2.add_all(k['eles/tgradcoru_upts'], deps=k['mpiint/con_u'])
        for l in k['eles/gradcoru_u']:
            g2.add(l, deps=deps(l, 'eles/tgradcoru_upts'))
        g2.commit()

        return g1, g2

    def compute_grads(self, t, uinbank):
        if uinbank not in self._ggraphs:
            self._ggraphs[uinbank] = self._compute_grads_graph(uinbank)

        for g in self._ggraphs[uinbank]:
            self.backend.run_graph(g)


system_map = {'euler': EulerSystem, 'navier-stokes': NavierStokesSystem}


def get_system(backend, mesh, cfg, nregs, comm=None, initsoln=None,
               needs_cfl=False):
    cls = system_map[cfg.get('solver', 'system')]
    sys = cls(backend, mesh, initsoln, nregs, cfg, comm, needs_cfl=needs_cfl)
    sys.commit()
    return sys
