#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the REFERENCE.

Runs only where /root/reference exists (the build container); the fixtures
it writes are committed so the CPU and GPU test suites can run anywhere.

    python tests/golden/make_golden.py

What is pinned, and by what:

* ``opmats.npz`` -- operator matrices produced by the reference's own
  ``pyfr/shapes.py`` (``BaseShape.opmat``, :79-135) for the element types
  and orders of BASELINE.json's configs.  The hex p=3 Gauss-Legendre
  ``m0..m3`` are additionally checked here against the reference's own
  golden file ``pyfr/tests/hex-gleg-ord3.npz`` (the single known-answer
  test the reference ships, ``pyfr/tests/test_ele_mats.py:10-27``).
* ``host_<case>.npz`` -- the reference's own ``NavierStokesSystem`` /
  ``EulerSystem`` (``pyfr/solvers/*``; imported with name-only stubs for
  the absent third-party packages, oracle/refharness.py) run on the NumPy
  oracle backend: every view index array the reference's host code hands
  to the backend (``View.mapping`` / ``rstrides``), the halo message
  sizes, the initial condition and the resulting RHS.  These pin this
  repository's host mirror (pyfr_b200/host): same mesh + same backend must
  give bit-identical indices and the same RHS to round-off.
"""

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refharness as rh                       # noqa: E402
from oracle.npbackend import LocalComm, make_backend      # noqa: E402
from pyfr_b200 import cases                               # noqa: E402

# name -> (case, mesh, case kwargs, partition bricks, oracle-backend options)
HOST_CASES = {
    'tgv_p2_rusanov': ('tgv', (3, 2, 2), dict(order=2, warp=0.1),
                       (1, 1, 1), {}),
    'tgv_p3_hllc_blocked': ('tgv', (3, 2, 2),
                            dict(order=3, rsolver='hllc', warp=0.1),
                            (1, 1, 1), {'blocks': 1, 'soasz': 8,
                                        'csubsz': 8}),
    'tgv_p2_beta0_2parts': ('tgv', (4, 2, 3), dict(order=2, beta=0.0,
                                                   warp=0.1),
                            (2, 1, 1), {}),
    'tgv_p1_4parts_blocked': ('tgv', (4, 4, 2), dict(order=1, warp=0.05),
                              (2, 2, 1), {'blocks': 1, 'soasz': 8,
                                          'csubsz': 8}),
    'tgv_p4_rusanov': ('tgv', (2, 2, 2), dict(order=4, warp=0.1),
                       (1, 1, 1), {}),
    'vortex_p3_rusanov': ('vortex', 4, dict(order=3), (1, 1), {}),
    'vortex_p3_hllc_2parts': ('vortex', (6, 4), dict(order=3,
                                                     rsolver='hllc'),
                              (2, 1), {'blocks': 1, 'soasz': 8,
                                       'csubsz': 16}),
    # flux anti-aliasing (quadrature-point flux, M7 / M9 operators)
    'tgv_p2_fluxaa': ('tgv', (3, 2, 2), dict(order=2, warp=0.1,
                                             antialias='flux'),
                      (1, 1, 1), {}),
    'tgv_p3_fluxaa_blocked_2parts': ('tgv', (4, 2, 2),
                                     dict(order=3, warp=0.1, rsolver='hllc',
                                          beta=0.0, antialias='flux'),
                                     (2, 1, 1), {'blocks': 1, 'soasz': 8,
                                                 'csubsz': 8}),
    'vortex_p3_fluxaa': ('vortex', 5, dict(order=3, antialias='flux'),
                         (1, 1), {}),
    # surface-flux anti-aliasing (flux points on a face quadrature rule,
    # projection folded into M3), alone and with the volume variant
    'tgv_p2_surfaa': ('tgv', (3, 2, 2), dict(order=2, warp=0.1,
                                             antialias='surf-flux'),
                      (1, 1, 1), {}),
    'vortex_p3_bothaa_2parts': ('vortex', (6, 4),
                                dict(order=3, rsolver='hllc',
                                     antialias='flux, surf-flux'),
                                (2, 1), {'blocks': 1, 'soasz': 8,
                                         'csubsz': 8}),
    # flux points coinciding with solution points (Gauss-Lobatto)
    'tgv_p3_gll_beta0_2parts': ('tgv', (4, 2, 2),
                                dict(order=3, warp=0.1, beta=0.0,
                                     pts='gauss-legendre-lobatto'),
                                (2, 1, 1), {}),
    'tgv_p2_gll_blocked': ('tgv', (3, 2, 2),
                           dict(order=2, warp=0.1,
                                pts='gauss-legendre-lobatto'),
                           (1, 1, 1), {'blocks': 1, 'soasz': 8, 'csubsz': 8}),
    'vortex_p3_gll': ('vortex', 5, dict(order=3,
                                        pts='gauss-legendre-lobatto'),
                      (1, 1), {}),
}

# Wall-bounded / open-boundary cases: name -> box_case arguments
BC_CASES = {
    'bc_ns_wall_farfield': ('navier-stokes', (3, 3, 2),
                            {'ylo': 'no-slp-adia-wall',
                             'yhi': 'char-riem-inv'},
                            dict(order=2, warp=0.1)),
    'bc_ns_inout_walls': ('navier-stokes', (3, 2, 3),
                          {'xlo': 'sub-in-frv', 'xhi': 'sub-out-fp',
                           'zlo': 'slp-adia-wall',
                           'zhi': 'no-slp-isot-wall'},
                          dict(order=2, rsolver='hllc')),
    'bc_ns_supersonic': ('navier-stokes', (2, 3, 3),
                         {'xlo': 'sup-in-fa', 'xhi': 'sup-out-fn'},
                         dict(order=1, beta=0.0)),
    'bc_ns_total_inflow': ('navier-stokes', (3, 2, 2),
                           {'xlo': 'sub-in-ftpttang', 'xhi': 'sub-out-fp'},
                           dict(order=2, warp=0.1)),
    'bc_euler_all': ('euler', (5, 4),
                     {'xlo': 'char-riem-inv', 'xhi': 'sup-out-fn',
                      'ylo': 'slp-adia-wall', 'yhi': 'sup-in-fa'},
                     dict(order=3)),
}

# Mixed element types (BASELINE.json configs[3]): name -> mixed_case arguments
MIXED_CASES = {
    'mixed_quad_tri_p3': ('quad+tri', (4, 3), dict(order=3, rsolver='hllc')),
    'mixed_hex_pri_p2': ('hex+pri', (3, 2, 2), dict(order=2, beta=0.0)),
    'mixed_all_p3': ('hex+pri+pyr+tet', (4, 2, 2), dict(order=3)),
}


def cat_fields(arrs):
    """One array per case: the field itself for a single element type,
    the element types' fields flattened one after the other otherwise."""
    return arrs[0] if len(arrs) == 1 else np.concatenate(
        [a.ravel() for a in arrs])


OPMAT_SHAPES = [('quad', 3, 'gauss-legendre'), ('hex', 2, 'gauss-legendre'),
                ('hex', 3, 'gauss-legendre'), ('hex', 4, 'gauss-legendre'),
                ('hex', 4, 'gauss-legendre-lobatto'),
                ('hex', 6, 'gauss-legendre')]
OPMAT_EXPRS = ['M0', 'M4 - M6*M0', 'M6', 'M1 - M3*M2', 'M3']
# anti-aliasing operators, for the smaller shapes only (they are dense)
OPMAT_AA_EXPRS = ['M7', '(M1 - M3*M2)*M9']
OPMAT_AA_SHAPES = [('quad', 3, 'gauss-legendre'), ('hex', 2, 'gauss-legendre'),
                   ('hex', 3, 'gauss-legendre')]


def cfg_text(case, kw, beopts):
    kw = {k: v for k, v in kw.items() if k != 'warp'}
    txt = cases.tgv_cfg(**kw) if case == 'tgv' else cases.vortex_cfg(**kw)
    if beopts:
        txt += '\n[backend-oracle]\n' + ''.join(f'{k} = {v}\n'
                                                for k, v in beopts.items())
    return txt


def record_views(be):
    """Wrap ``be.view`` so every view's index arrays are recorded."""
    trace, orig = [], be.view

    def view(*a, **k):
        v = orig(*a, **k)
        trace.append(v)
        return v

    be.view = view
    return trace


def record_consts(becls):
    """Wrap ``const_matrix`` of a backend base class so the floating-point
    constant tables (normals, metric terms, vertices) are recorded.
    Returns (list, undo callable)."""
    rec, orig = [], becls.const_matrix

    def const_matrix(self, initval, *a, **k):
        iv = np.asarray(initval)
        tags = k.get('tags', a[1] if len(a) > 1 else set())
        if iv.dtype.kind == 'f' and not any(t.startswith('M') for t in tags):
            rec.append(np.array(iv))
        return orig(self, initval, *a, **k)

    becls.const_matrix = const_matrix
    return rec, lambda: setattr(becls, 'const_matrix', orig)


def consts_digest(rec):
    """Constant tables in a creation-order independent order."""
    def key(a):
        w = np.arange(1, a.size + 1).reshape(a.shape)/a.size
        return (a.shape, round(float(np.abs(a).sum()), 6),
                round(float(a.sum()), 6), round(float((a*w).sum()), 6))

    return sorted(rec, key=key)


def trace_digest(trace):
    """Order-independent fingerprint + the raw arrays of a view trace."""
    arrs = []
    for v in trace:
        arrs.append(np.ascontiguousarray(v.mapping.get()[0]))
        if v.rstrides is not None:
            arrs.append(np.ascontiguousarray(v.rstrides.get()[0]))

    keys = sorted(f'{a.dtype}:{a.shape}:' + hashlib.sha256(a.tobytes())
                  .hexdigest() for a in arrs)
    return keys, arrs


def ref_opmats():
    rh.install_stubs()
    from pyfr.inifile import Inifile
    from pyfr.shapes import HexShape, QuadShape

    out = {}
    for et, order, pts in OPMAT_SHAPES:
        face = 'line' if et == 'quad' else 'quad'
        cfg = Inifile(f'[solver]\norder = {order}\n'
                      f'[solver-elements-{et}]\nsoln-pts = {pts}\n'
                      f'[solver-interfaces-{face}]\nflux-pts = {pts}\n')
        shape = {'quad': QuadShape, 'hex': HexShape}[et](None, cfg)

        exprs = OPMAT_EXPRS + (OPMAT_AA_EXPRS
                               if (et, order, pts) in OPMAT_AA_SHAPES else [])
        for expr in exprs:
            out[f'{et}|{order}|{pts}|{expr}'] = shape.opmat(expr)

        if (et, order, pts) == ('hex', 3, 'gauss-legendre'):
            kat = np.load('/root/reference/pyfr/tests/hex-gleg-ord3.npz')
            for m in ('m0', 'm1', 'm2', 'm3'):
                assert np.allclose(getattr(shape, m), kat[m]), m
                out[f'kat|hex|3|gauss-legendre|{m}'] = getattr(shape, m)

    return out


def ref_host_case(name):
    if name in BC_CASES:
        system, n, bcs, kw = BC_CASES[name]
        _, box, txt = cases.box_case(system, n, bcs, **kw)
        parts = (1,)*box.ndims
    elif name in MIXED_CASES:
        pattern, n, kw = MIXED_CASES[name]
        _, box, txt = cases.mixed_case(pattern, n, **kw)
        parts = (1,)*box.ndims
    else:
        case, n, kw, parts, beopts = HOST_CASES[name]
        txt = cfg_text(case, kw, beopts)
        _, box = cases.make(case, n, **kw)
    nparts = int(np.prod(parts))
    vparts = box.brick_partition(parts) if nparts > 1 else None

    world = LocalComm(0, nparts)
    systems, traces, consts = [], [], []

    rh.install_stubs()
    import pyfr.backends.base.backend as rbb

    for r in range(nparts):
        # Record the views the reference's host code creates
        holder = {}
        orig_init = rbb.BaseBackend.__init__

        def init(self, cfg, _h=holder, _o=orig_init):
            _o(self, cfg)
            _h['trace'] = record_views(self)

        rbb.BaseBackend.__init__ = init
        crec, undo = record_consts(rbb.BaseBackend)
        try:
            s, be = rh.ref_system(txt, box.local_mesh(vparts, r), 2,
                                  world.peer(r))
        finally:
            rbb.BaseBackend.__init__ = orig_init
            undo()

        consts.append(consts_digest(crec))

        systems.append(s)
        traces.append(holder['trace'])

    graphs = [s._rhs_graphs(0, 1) for s in systems]
    for s in systems:
        s._prepare_kernels(0.0, 0, 1)
    for stage in zip(*graphs):
        for g in stage:
            g.run()
        world.deliver()

    out = {}
    for r, (s, tr) in enumerate(zip(systems, traces)):
        keys, arrs = trace_digest(tr)
        out[f'r{r}_viewkeys'] = np.array(keys)
        for i, a in enumerate(arrs):
            out[f'r{r}_view{i}'] = a
        for i, a in enumerate(consts[r]):
            out[f'r{r}_const{i}'] = a
        out[f'r{r}_ics'] = cat_fields(s.ele_scal_upts(0))
        out[f'r{r}_rhs'] = cat_fields(s.ele_scal_upts(1))

    return out


# Connectivity cases: name -> (mesh size, bounds, periodic axes, nparts)
CONN_CASES = {
    'conn_hex_periodic_3parts': ((5, 4, 3), (-1.0, 1.0), True, 3),
    'conn_hex_walls_4parts': ((4, 4, 4), (0.0, 2.0), (True, False, False), 4),
    'conn_quad_walls_5parts': ((9, 7), (0.0, 1.0), (False, True), 5),
}


PARTITION_SIZES = (6, 64, 128)


def ref_partitions(n, nparts=(2, 4, 8)):
    """``vparts<N>``: the partition of a periodic ``n``^3 hex box into N
    parts by the reference's ``BaselinePartitioner`` (default options, unit
    weights; ``pyfr/partitioners/baseline.py:87-105`` on the dual graph the
    reference builds in ``pyfr/partitioners/base.py``)."""
    rh.install_stubs()
    from pyfr.partitioners.base import Graph
    from pyfr.partitioners.baseline import BaselinePartitioner

    from pyfr_b200.host.mesh import BoxMesh

    box = BoxMesh((n, n, n), 0.0, 1.0, periodic=True)
    ne = box.neles

    # Dual graph: every element of a periodic box of n >= 3 has six
    # distinct neighbours
    nb = np.sort(box.roff, axis=1)
    assert (nb >= 0).all() and (np.diff(nb, axis=1) > 0).all()
    vtab = 6*np.arange(ne + 1, dtype=np.int64)
    etab = nb.ravel().astype(np.int64)

    out = {}
    for k in nparts:
        graph = Graph(vtab, etab, np.ones(ne, dtype=np.int32),
                      np.ones(len(etab), dtype=np.int32))
        part = BaselinePartitioner([1]*k, elewts={box.etype: 1})
        vp = np.asarray(part._partition_graph(graph, [1.0]*k))
        assert len(np.unique(vp)) == k
        out[f'vparts{k}'] = vp.astype(np.int8)

    return out


def ref_connectivity_case(name):
    """Partition a synthetic box with the reference's own
    ``BaselinePartitioner`` (pyfr/partitioners/baseline.py) and derive each
    rank's interior / boundary / inter-partition connectivity with the
    reference's own ``NativeReader._construct_con`` (pyfr/readers/native.py:
    445-534), every rank in a thread with an in-process stand-in for the
    MPI neighbourhood collectives.  What is fed in is the mesh in the form
    the reader would find in a ``.pyfrm`` file: per element, per face, the
    ``(codec index, global element)`` of the neighbour."""
    import threading
    from types import SimpleNamespace

    rh.install_stubs()
    import pyfr.readers.native as rnative
    from pyfr.partitioners.base import Graph
    from pyfr.partitioners.baseline import BaselinePartitioner

    from pyfr_b200.host.mesh import BoxMesh

    n, (lo, hi), periodic, nparts = CONN_CASES[name]
    box = BoxMesh(n, lo, hi, periodic=periodic)
    et, ne, nfaces = box.etype, box.neles, box.roff.shape[1]

    # Dual graph (unit weights) -> the reference's partitioner
    nbr = [np.unique(r[(r >= 0) & (r != g)]) for g, r in enumerate(box.roff)]
    vtab = np.concatenate(([0], np.cumsum([len(x) for x in nbr])))
    etab = np.concatenate(nbr)
    graph = Graph(vtab, etab, np.ones(ne, dtype=np.int32),
                  np.ones(len(etab), dtype=np.int32))
    part = BaselinePartitioner([1]*nparts, elewts={et: 1})
    vparts = np.asarray(part._partition_graph(graph, [1.0]*nparts),
                        dtype=np.int32)
    assert len(np.unique(vparts)) == nparts

    order = box.partition_order(vparts)

    # In-process neighbourhood collectives
    tls = threading.local()
    barrier = threading.Barrier(nparts)
    mail = {}

    class NComm:
        handle = 0

        def __init__(self, nbrs):
            self.nbrs = nbrs

        @staticmethod
        def fromhandle(h):
            return SimpleNamespace(free=lambda: None)

        def neighbor_allgather(self, obj):
            mail['g', tls.rank] = obj
            barrier.wait()
            out = [mail['g', p] for p in self.nbrs]
            barrier.wait()
            return out

        def neighbor_alltoall(self, objs):
            for p, o in zip(self.nbrs, objs):
                mail['a', tls.rank, p] = o
            barrier.wait()
            out = [mail['a', p, tls.rank] for p in self.nbrs]
            barrier.wait()
            return out

    class Comm:
        def Create_dist_graph_adjacent(self, src, dst):
            return NComm(list(src))

    rnative.get_comm_rank_root = lambda: (Comm(), tls.rank, 0)

    results, errors = {}, []

    def run(rank):
        try:
            tls.rank = rank
            gidx = order[rank]
            faces = np.empty((len(gidx), nfaces),
                             dtype=[('cidx', np.int16), ('off', np.int64)])
            faces['cidx'], faces['off'] = box.rcidx[gidx], box.roff[gidx]

            rd = rnative.NativeReader.__new__(rnative.NativeReader)
            rd.mesh = SimpleNamespace(codec=list(box.codec), etypes=[et],
                                      eidxs={et: gidx}, bcon={}, con_p={})
            rd.eles = {et: {'faces': faces}}
            rd.f = {f'eles/{et}': np.empty(ne)}

            nb = vparts[box.roff[gidx][box.roff[gidx] >= 0]]
            rd.neighbours = sorted(set(nb.tolist()) - {rank})
            rd._construct_con()
            results[rank] = rd.mesh
        except Exception as e:                          # pragma: no cover
            errors.append(e)
            barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(nparts)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]

    out = {'vparts': vparts}
    for r, m in results.items():
        out[f'r{r}_eidxs'] = order[r]
        for side, c in zip('lr', m.con):
            out[f'r{r}_con_{side}_cidxs'] = c.cidxs
            out[f'r{r}_con_{side}_eidxs'] = c.eidxs
        for p, c in m.con_p.items():
            out[f'r{r}_conp{p}_cidxs'] = c.cidxs
            out[f'r{r}_conp{p}_eidxs'] = c.eidxs
        for b, c in m.bcon.items():
            out[f'r{r}_bcon_{b}_cidxs'] = c.cidxs
            out[f'r{r}_bcon_{b}_eidxs'] = c.eidxs

    return out


# Time integration: the reference's own integrator classes (stepper +
# controller composed by pyfr.integrators.get_integrator) on the oracle
# backend.  name -> (case, n, kw, [solver-time-integrator] options, target
# times handed to advance_to one after the other)
INTG_CASES = {
    'vortex_p3_rk45_pi_l2': (
        'vortex', (4, 4), dict(order=3),
        dict(scheme='rk45', controller='pi', dt=0.05, atol=1e-6, rtol=1e-6),
        [0.2, 0.33]),
    'tgv_p2_rk45_pi_uniform': (
        'tgv', (2, 2, 2), dict(order=2),
        {'scheme': 'rk45', 'controller': 'pi', 'dt': 0.02,
         'errest-norm': 'uniform', 'rtol': 1e-5, 'atol-rho': 1e-5,
         'atol-rhou': 2e-5, 'atol-rhov': 2e-5, 'atol-rhow': 2e-5,
         'atol-E': 1e-4, 'pi-alpha': 0.7, 'pi-beta': 0.4,
         'safety-fact': 0.9, 'max-fact': 1.5, 'min-fact': 0.5,
         'dt-lookahead': 3},
        [0.15]),
    'vortex_p3_rk45_none': (
        'vortex', (4, 4), dict(order=3),
        dict(scheme='rk45', controller='none', dt=0.01), [0.05]),
    'vortex_p3_rk4_none': (
        'vortex', (4, 4), dict(order=3),
        dict(scheme='rk4', controller='none', dt=0.01), [0.05]),
    'vortex_p3_rk4_cfl': (
        'vortex', (4, 4), dict(order=3),
        {'scheme': 'rk4', 'controller': 'cfl', 'dt': 0.01, 'cfl': 0.4,
         'cfl-nsteps': 2}, [0.12]),
    'tgv_p2_rk45_cfl_curved': (
        'tgv', (2, 2, 2), dict(order=2, warp=0.1),
        {'scheme': 'rk45', 'controller': 'cfl', 'dt': 0.01, 'cfl': 0.3,
         'dt-max': 0.02}, [0.1]),
}


def intg_cfg_text(name):
    case, n, kw, opts, tlist = INTG_CASES[name]
    sect = '\n'.join(f'{k} = {v}' for k, v in opts.items())
    return (cfg_text(case, kw, {}) + '\n[solver-time-integrator]\n'
            f'formulation = explicit\ntstart = 0\ntend = {tlist[-1]!r}\n'
            f'{sect}\n')


def ref_intg_case(name):
    case, n, kw, opts, tlist = INTG_CASES[name]
    _, box = cases.make(case, n, **kw)

    rh.install_stubs()
    rh.set_rank(LocalComm(0, 1).peer(0))

    import pyfr.backends.base as rbase
    from pyfr.inifile import Inifile
    from pyfr.integrators import get_integrator
    from pyfr.solvers.euler import EulerSystem
    from pyfr.solvers.navstokes import NavierStokesSystem

    cfg = Inifile(intg_cfg_text(name))
    be = make_backend(rbase, name='oracle-ref')(cfg)
    cls = {'euler': EulerSystem,
           'navier-stokes': NavierStokesSystem}[cfg.get('solver', 'system')]
    intg = get_integrator(be, cls, rh.ref_mesh(box.local_mesh()), None, cfg)

    # Record every accept / reject decision
    hist = []
    for what in ('accept', 'reject'):
        def wrap(dt, idx, wtime, err=None, _w=what,
                 _o=getattr(intg, f'_{what}_step')):
            hist.append((dt, _w == 'accept', -1.0 if err is None else err))
            _o(dt, idx, wtime, err=err)
        setattr(intg, f'_{what}_step', wrap)

    out = {'u0': intg.soln[0].copy()}
    for i, t in enumerate(tlist):
        intg.advance_to(t)
        out[f'u_t{i}'] = intg.soln[0].copy()
        out[f'tcurr_t{i}'] = np.array(intg.tcurr)

    out['hist'] = np.array(hist, dtype=float)
    out['counts'] = np.array([intg.nacptsteps, intg.nrjctsteps,
                              intg.nrhsevals, intg.gndofs])
    out['dt_final'] = np.array(intg.dt)
    return out


# Element types whose bases and point sets the host mirror does not
# construct itself: everything pyfr_b200/host needs of a shape, tabulated
# from the reference's shapes.py / polys.py / quadrules at the point sets of
# BASELINE.json configs[3].  Lives inside the
# package because the GPU box has no /root/reference.
TAB_SHAPES = [(et, p) for et in ('tri', 'tet', 'pri', 'pyr')
              for p in (1, 2, 3)]

TAB_POINTS = {
    'tri': 'williams-shunn', 'tet': 'shunn-ham',
    'pri': 'williams-shunn~gauss-legendre', 'pyr': 'gauss-legendre',
    'line': 'gauss-legendre', 'quad': 'gauss-legendre',
}


def ref_tabulated_shapes():
    from pyfr.inifile import Inifile
    from pyfr.quadrules import get_quadrule
    from pyfr.shapes import BaseShape
    from pyfr.util import subclass_where

    out = {}
    for et, order in TAB_SHAPES:
        cfg = Inifile(
            f'[solver]\norder = {order}\n' +
            ''.join(f'[solver-elements-{k}]\nsoln-pts = {v}\n'
                    f'[solver-interfaces-{k}]\nflux-pts = {v}\n'
                    for k, v in TAB_POINTS.items())
        )
        scls = subclass_where(BaseShape, name=et)
        nverts = len(scls.std_ele(1))
        sh = scls(nverts, cfg)
        pre = f'{et}_p{order}_'

        named = {'upts': sh.upts, 'fpts': sh.fpts, 'mpts': sh.mpts,
                 'linspts': sh.linspts}
        for n, pts in named.items():
            pts = np.asarray(pts, dtype=float)
            out[pre + n] = pts
            out[pre + f'sbasis@{n}'] = sh.sbasis.nodal_basis_at(pts)
            out[pre + f'mbasis@{n}'] = sh.mbasis.nodal_basis_at(pts)
        J = sh.mbasis.jac_nodal_basis_at(sh.mpts)       # (nd, nbasis, npts)
        out[pre + 'mbasis_deriv@mpts'] = np.array([Jd.T for Jd in J])

        rname = cfg.get(f'solver-elements-{et}', 'soln-pts')
        out[pre + 'upts_wts'] = get_quadrule(et, rname, sh.nupts).wts
        out[pre + 'norm_fpts'] = sh.norm_fpts
        out[pre + 'nfacefpts'] = np.array(sh.nfacefpts)
        out[pre + 'facefpts'] = np.concatenate(
            [np.asarray(f) for f in sh.facefpts])
        out[pre + 'jac_exprs'] = np.array(sh.jac_exprs)
        for m in ('m0', 'm1', 'm2', 'm3', 'm4', 'm6'):
            out[pre + m] = getattr(sh, m)

        # Vertices on each face (face pairing of linear meshes)
        fverts, counts = [], []
        fc = {'line': [(-1,), (1,)], 'quad': [(-1, -1), (1, -1), (-1, 1),
                                              (1, 1)],
              'tri': [(-1, -1), (1, -1), (-1, 1)]}
        lin = np.asarray(sh.linspts, dtype=float)
        for ftype, proj, _ in scls.faces:
            ids = [int(np.argmin(np.abs(lin - np.array(proj(*c),
                                                       dtype=float)).sum(1)))
                   for c in fc[ftype]]
            assert all(np.allclose(lin[i], proj(*c))
                       for i, c in zip(ids, fc[ftype]))
            fverts += ids
            counts.append(len(ids))
        out[pre + 'faceverts'] = np.array(fverts)
        out[pre + 'nfaceverts'] = np.array(counts)

    return out


def main():
    rh.install_stubs()

    if sys.argv[1:] in ([], ['--shapes']):
        path = os.path.join(ROOT, 'pyfr_b200', 'host', 'data',
                            'tabshapes.npz')
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez_compressed(path, **ref_tabulated_shapes())
        print('pyfr_b200/host/data/tabshapes.npz written')

        if sys.argv[1:]:
            return

    if sys.argv[1:] in ([], ['--intg']):
        for name in INTG_CASES:
            np.savez_compressed(os.path.join(HERE, f'intg_{name}.npz'),
                                **ref_intg_case(name))
            print(f'intg_{name}.npz written')

        if sys.argv[1:]:
            return

    if sys.argv[1:] in ([], ['--kernels']):
        # Input/output vectors of the reference's kernel templates, rendered
        # by oracle/minimako.py and compiled as C: recorded while
        # tests/test_oracle_templates.py runs against the live templates
        import subprocess

        path = os.path.join(HERE, 'kernel_vectors.npz')
        subprocess.run(
            [sys.executable, '-m', 'pytest', '-q', '-x',
             os.path.join(ROOT, 'tests', 'test_oracle_templates.py')],
            env=dict(os.environ, PYFR_B200_RECORD_KERNELS=path), check=True,
            cwd=ROOT
        )
        print('kernel_vectors.npz written')

        if sys.argv[1:]:
            return

    if sys.argv[1:2] == ['--only']:
        for name in sys.argv[2].split(','):
            np.savez_compressed(os.path.join(HERE, f'host_{name}.npz'),
                                **ref_host_case(name))
            print(f'host_{name}.npz written')
        return

    if sys.argv[1:2] == ['--partitions']:
        # parts_hex<n>.npz: the reference partitioner's element -> rank
        # maps of periodic n^3 hex boxes (what bench.py --partition
        # reference and the strong-scaling runs use)
        for n in map(int, sys.argv[2].split(',') if sys.argv[2:] else
                     PARTITION_SIZES):
            np.savez_compressed(os.path.join(HERE, f'parts_hex{n}.npz'),
                                **ref_partitions(n))
            print(f'parts_hex{n}.npz written')
        return

    if sys.argv[1:] == ['--mixed']:
        for name in MIXED_CASES:
            np.savez_compressed(os.path.join(HERE, f'host_{name}.npz'),
                                **ref_host_case(name))
            print(f'host_{name}.npz written')
        return

    for name in CONN_CASES:
        np.savez_compressed(os.path.join(HERE, f'{name}.npz'),
                            **ref_connectivity_case(name))
        print(f'{name}.npz written')

    np.savez_compressed(os.path.join(HERE, 'opmats.npz'), **ref_opmats())
    print('opmats.npz written')

    for name in list(HOST_CASES) + list(BC_CASES) + list(MIXED_CASES):
        np.savez_compressed(os.path.join(HERE, f'host_{name}.npz'),
                            **ref_host_case(name))
        print(f'host_{name}.npz written')


if __name__ == '__main__':
    main()
