"""CPU tests of the oracle physics (analytic checks), of the host logic
(partition invariance, layout, view arithmetic) and of what the B200
backend *plans* to launch (dry runtime: kernels are generated and the graph
rewrites applied, nothing is executed)."""

import re

import numpy as np
import pytest

from oracle.npbackend import LocalComm
from pyfr_b200 import base, cases
from pyfr_b200.host.config import Config
from pyfr_b200.host.system import get_system

from util import (OracleBackend, conservation_defect, oracle_rhs, rel_err,
                  run_lockstep)


def _with_ics(txt, ics):
    return txt[:txt.index('[soln-ics]')] + '[soln-ics]\n' + ics


def _oracle_system(cfg, box, **beopts):
    for k, v in beopts.items():
        cfg.set('backend-oracle', k, v)
    be = OracleBackend(cfg)
    return get_system(be, box.local_mesh(), cfg, 2)


# -- analytic checks of the restated kernel arithmetic ---------------------
@pytest.mark.parametrize('case,n,kw', [
    ('tgv', (3, 3, 2), dict(order=3, warp=0.15)),
    ('tgv', (3, 2, 2), dict(order=2, warp=0.1, rsolver='hllc', beta=0.0)),
    ('vortex', 5, dict(order=3)),
])
def test_free_stream_preserved(case, n, kw):
    """Uniform flow on a non-affine mesh has zero RHS (metric identities,
    consistency of the Riemann solvers and of the LDG fluxes)."""
    warp = kw.pop('warp', None)
    txt = cases.tgv_cfg(**kw) if case == 'tgv' else cases.vortex_cfg(**kw)
    w = 'w = 0.1\n' if case == 'tgv' else ''
    cfg = Config(_with_ics(txt, f'rho = 1.2\nu = 0.3\nv = -0.2\n{w}p = 2.5\n'))
    box = (cases.tgv_mesh(n, warp=warp) if case == 'tgv'
           else cases.vortex_mesh(n))

    s = _oracle_system(cfg, box)
    s.rhs(0.0, 0, 1)

    assert np.abs(s.ele_scal_upts(1)[0]).max() < 2e-12


def test_vortex_rhs_converges_at_design_order():
    """The isentropic vortex translates with velocity (0, 1): the exact
    time derivative is -d/dy of the initial condition.  The Euler RHS
    must converge to it at order >= p between two meshes."""
    errs = []
    for n in (40, 80):
        cfg, box = cases.make('vortex', n, order=3)
        s = _oracle_system(cfg, box)
        s.rhs(0.0, 0, 1)
        rhs = s.ele_scal_upts(1)[0]

        # Central difference of the analytic IC in y (h^2 error 1e-10)
        h = 1e-4
        ics = []
        for sgn in (+1, -1):
            txt = cases.vortex_cfg(order=3)
            i = txt.index('[soln-ics]')
            txt = txt[:i] + re.sub(r'\by\b', f'(y + {sgn*h})', txt[i:])
            c2 = Config(txt)
            s2 = _oracle_system(c2, box)
            ics.append(s2.ele_scal_upts(0)[0])
        exact = -(ics[0] - ics[1])/(2*h)

        errs.append(np.sqrt(np.mean((rhs - exact)**2)))

    assert errs[1] < errs[0]/8.0, errs


def test_shear_layer_viscous_rhs_converges_to_exact():
    """rho, p uniform, u = A sin(y), v = w = 0: convection and pressure
    terms vanish identically and d(rho u)/dt = mu u'' = -mu A sin(y),
    dE/dt = d/dy(mu u u') = mu A^2 cos(2y).  Exercises the whole viscous
    path (gradient correction, LDG common solution/flux, divergence): the
    discrete RHS must converge to the exact one at high order."""
    from pyfr_b200.host.elements import NavierStokesElements
    from pyfr_b200.host.shapes import shape_map

    A, mu, errs = 0.5, 0.05, []
    ics = f'rho = 1\nu = {A}*sin(y)\nv = 0\nw = 0\np = 10\n'
    txt = cases.tgv_cfg(order=4).replace('mu = 6.25e-4', f'mu = {mu}')

    for n in (8, 16):
        cfg = Config(_with_ics(txt, ics))
        box = cases.tgv_mesh((2, n, 2))
        s = _oracle_system(cfg, box)
        s.rhs(0.0, 0, 1)
        rhs = s.ele_scal_upts(1)[0]

        e = NavierStokesElements(shape_map['hex'],
                                 box.local_mesh().spts['hex'], cfg)
        y = e.ploc_at_np('upts')[:, 1]

        errs.append((np.abs(rhs[:, 1] + mu*A*np.sin(y)).max()/(mu*A),
                     np.abs(rhs[:, 4] - mu*A*A*np.cos(2*y)).max()/(mu*A*A)))
        assert np.abs(rhs[:, 0]).max() < 1e-12
        assert np.abs(rhs[:, 2]).max() < 1e-4*(8/n)**4

    (m0, e0), (m1, e1) = errs
    assert m1 < 2e-3 and m1 < m0/8, errs
    assert e1 < 2e-2 and e1 < e0/6, errs


# -- partition invariance -------------------------------------------------------
@pytest.mark.parametrize('case,n,parts,kw', [
    ('tgv', (4, 2, 2), (2, 1, 1), dict(order=2, warp=0.1, beta=0.0)),
    ('tgv', (4, 4, 2), (2, 2, 1), dict(order=2, warp=0.1, beta=0.0,
                                       rsolver='hllc')),
    ('tgv', (4, 4, 4), (2, 2, 2), dict(order=1, warp=0.05, beta=0.0)),
    ('vortex', (6, 4), (2, 1), dict(order=3)),
    ('vortex', (4, 6), (2, 2), dict(order=2, rsolver='hllc')),
])
def test_partitioned_rhs_equals_single_partition(case, n, parts, kw):
    """N partitions exchanging halos (in-process ranks) reproduce the
    single-partition RHS element for element.  (Only for Euler and for
    LDG beta = 0: with beta != 0 the reference orients the one-sided LDG
    fluxes of an inter-partition face by rank parity, pyfr/solvers/
    baseadvecdiff/inters.py:47-58, and those of an interior face by face
    order, so a partitioned run is a *different*, equally valid
    discretisation; that case is pinned by the reference-host fixtures in
    test_oracle_golden.py and by conservation below.)"""
    warp = {'warp': kw.pop('warp')} if 'warp' in kw else {}
    _, ref = oracle_rhs(case, n, **warp, **kw)

    _, box = cases.make(case, n, **warp, **kw)
    nparts = int(np.prod(parts))
    vparts = box.brick_partition(parts)
    systems, out = oracle_rhs(case, n, vparts=vparts, nparts=nparts, **warp,
                              **kw)

    et = box.etype
    seen = np.zeros(box.neles, dtype=int)
    for r, (s, o) in enumerate(zip(systems, out)):
        gidx = box.local_mesh(vparts, r).eidxs[et]
        seen[gidx] += 1
        assert rel_err(o, ref[0][..., gidx]) < 5e-13

    assert np.all(seen == 1)


@pytest.mark.parametrize('case,n,parts,kw', [
    ('tgv', (3, 2, 2), (1, 1, 1), dict(order=3, warp=0.1)),
    ('tgv', (4, 2, 2), (2, 1, 1), dict(order=2, warp=0.1, rsolver='hllc')),
    ('tgv', (2, 2, 4), (1, 1, 2), dict(order=2, beta=-0.5)),
    ('tgv', (4, 4, 2), (2, 2, 1), dict(order=1, warp=0.1, beta=0.25)),
    ('vortex', (6, 4), (2, 1), dict(order=3)),
])
def test_rhs_is_conservative(case, n, parts, kw):
    """The RHS integrates to zero over the periodic domain for every
    conserved variable, for any partitioning and LDG orientation."""
    warp = {'warp': kw.pop('warp')} if 'warp' in kw else {}
    cfg, box = cases.make(case, n, **warp, **kw)
    nparts = int(np.prod(parts))
    vparts = box.brick_partition(parts) if nparts > 1 else None
    systems, out = oracle_rhs(case, n, vparts=vparts, nparts=nparts, **warp,
                              **kw)

    tot = mag = 0
    for r, o in enumerate(out):
        t, m = conservation_defect(cfg, box.local_mesh(vparts, r), o)
        tot, mag = tot + t, mag + m

    assert np.all(np.abs(tot) <= 1e-12*mag.max()), (tot, mag)


def test_mpi_interfaces_agree_on_point_order():
    """Both sides of an inter-partition interface list their flux points
    in the same physical order (what makes the dense halo message
    meaningful without an index exchange)."""
    from pyfr_b200.host.elements import NavierStokesElements
    from pyfr_b200.host.shapes import shape_map

    cfg, box = cases.make('tgv', (4, 2, 2), order=2, warp=0.1)
    vparts = box.brick_partition((2, 1, 1))
    L = box.hi - box.lo

    locs = {}
    for r in range(2):
        m = box.local_mesh(vparts, r)
        e = NavierStokesElements(shape_map['hex'], m.spts['hex'], cfg)
        con = m.con_p[1 - r]
        pts = []
        for etype, fidx, eidxs, _ in con.foreach():
            rows = e.srtd_face_fpts[fidx][eidxs]
            pts.append((con, fidx, e.plocfpts[rows, eidxs[:, None]]))
        # Re-assemble in connectivity order
        full = np.empty((len(con), e.nfacefpts[0], 3))
        for etype, fidx, eidxs, ix in con.foreach():
            rows = e.srtd_face_fpts[fidx][eidxs]
            full[ix] = e.plocfpts[rows, eidxs[:, None]]
        locs[r] = full

    d = np.abs(locs[0] - locs[1])
    d = np.minimum(d, np.abs(d - L))          # periodic images coincide
    assert d.max() < 1e-9


# -- storage layout and view arithmetic ---------------------
@pytest.mark.parametrize('blocks,k,csub', [(0, 8, 8), (1, 8, 8), (1, 4, 8),
                                           (1, 8, 16)])
def test_matrix_layout_formula(blocks, k, csub):
    cfg, _ = cases.make('tgv', 2)
    for o, v in (('blocks', blocks), ('soasz', k), ('csubsz', csub)):
        cfg.set('backend-oracle', o, v)
    be = OracleBackend(cfg)

    nrow, nvar, n = 7, 5, 37
    a = np.random.default_rng(1).standard_normal((nrow, nvar, n))
    m = be.matrix((nrow, nvar, n), a, tags={'align'})
    be.commit()

    assert np.array_equal(m.get(), a)

    flat = m.basedata[m.offset:m.offset + m.nbytes].view(np.float64)
    for (r, v, e) in [(0, 0, 0), (3, 2, 17), (6, 4, 36), (5, 1, 8)]:
        if blocks:
            # [n/csub][nrow][csub/k][nvar][k]
            b, ee = divmod(e, csub)
            ix = (b*m.blocksz + r*m.leaddim + (ee // k)*nvar*k + v*k + ee % k)
        else:
            # [nrow][n/k][nvar][k]
            ix = r*m.leaddim + (e // k)*nvar*k + v*k + e % k
        assert flat[ix] == a[r, v, e]


def test_view_mapping_formula():
    """mapping = offset + block displacement + row*leaddim + SoA column
    (pyfr/backends/base/types.py:294-320)."""
    cfg, _ = cases.make('tgv', 2)
    cfg.set('backend-oracle', 'blocks', 1)
    be = OracleBackend(cfg)

    nrow, nvar, n, k = 6, 5, 20, be.soasz
    m = be.matrix((nrow, nvar, n), tags={'align'}, extent='x')
    m2 = be.matrix((nrow, nvar, n), tags={'align'}, extent='x')
    be.commit()

    rng = np.random.default_rng(2)
    rmap, cmap = rng.integers(0, nrow, 50), rng.integers(0, n, 50)
    mats = np.where(rng.integers(0, 2, 50) == 0, m.mid, m2.mid)
    v = be.view(mats, rmap, cmap, vshape=(nvar,))
    be.commit()

    got = v.mapping.get()[0]
    it = m.itemsize
    for i in range(50):
        mm = m if mats[i] == m.mid else m2
        b, e = divmod(cmap[i], be.csubsz)
        exp = (mm.offset//it + b*mm.blocksz + rmap[i]*mm.leaddim +
               (e // k)*nvar*k + e % k)
        assert got[i] == exp


# -- what the B200 backend plans (no device needed) ---------------------------
def _dry_plan(case, n, opts={}, nparts=1, **kw):
    from pyfr_b200.backend import B200Backend

    cfg, box = cases.make(case, n, **kw)
    for k, v in opts.items():
        cfg.set('backend-b200', k, v)

    be = B200Backend(cfg, dry=True)
    parts = (2,) + (1,)*(box.ndims - 1)
    vparts = box.brick_partition(parts) if nparts > 1 else None
    comm = type('C', (), dict(rank=0, size=nparts))()
    s = get_system(be, box.local_mesh(vparts, 0), cfg, 2, comm=comm)

    return be, [[(w, getattr(k, 'kind', None) if w == 'kernel' else
                  [r.kind for r in k]) for w, k in g.plan]
                for g in s.rhs_graphs(0, 1)]


def test_b200_ns_rhs_is_four_launches(built):
    """disu; the element kernel (which gathers the common solution itself:
    the interior intconu of the first graph is folded into it); the common
    flux; the correction with negdivconf."""
    be, plan = _dry_plan('tgv', 2, order=4)
    kinds = [[k for w, k in g] for g in plan]

    assert kinds == [['mul'], ['gradflux', None], ['mul+negdivconf']]

    # without the fold: five, intconu storing both sides (no copy_fpts)
    be, plan = _dry_plan('tgv', 2, {'conu-fold': 0}, order=4)
    kinds = [[k for w, k in g] for g in plan]

    assert kinds == [['mul', 'intconu'], ['gradflux', None],
                     ['mul+negdivconf']]

    # a central LDG flux averages the two traces: nothing to fold
    be, plan = _dry_plan('tgv', 2, order=4, beta=0.0)
    assert [k for w, k in plan[0]] == ['mul', 'intconu']


def test_conu_fold_gather_indices(built):
    """The index table the folded element kernel gathers through: every
    flux point of every element names the entry of scal_fpts that the
    reference's intconu would have copied into it -- for ldg-beta = +1/2
    the right-hand trace of its interface."""
    from pyfr_b200.backend import B200Backend

    for beta, rows in ((0.5, 0), (-0.5, 0), (0.5, 1)):
        cfg, box = cases.make('tgv', (3, 2, 2), order=2, beta=beta)
        cfg.set('backend-b200', 'gather-rows', rows)
        be = B200Backend(cfg, dry=True)
        s = get_system(be, box.local_mesh(), cfg, 2)
        g0, g1, g2 = s.rhs_graphs(0, 1)

        conu, = g0.folded
        gf, = [k for w, k in g1.plan if w == 'kernel' and k.kind == 'gradflux']
        assert gf.info['gather'] and conu.kind == 'intconu'

        i = conu.info
        S = i['ulin']._mats[0]
        csub, nf, nb = be.csubsz, S.nrow, S.nblocks
        gidx = gf.info['gidx'].get().reshape(nb, nf, csub)
        src = (i['urin'] if beta > 0 else i['ulin']).mapping.get()[0]
        src = src - S.offset // S.itemsize

        if rows:
            # Whole rows (their points marked -2) go by one bulk copy from
            # rowd[block, row]: put the points back and the table must be
            # the per-point one
            rowd = gf.info['rowd'].get().reshape(nb, nf + 1)
            whole = rowd[:, :nf] >= 0
            assert whole.any() and not whole.all()
            assert np.array_equal(whole.sum(axis=1), rowd[:, nf])
            assert np.array_equal((gidx == -2).all(axis=2), whole)
            assert np.array_equal((gidx == -2).any(axis=2), whole)
            assert (rowd[:, :nf][whole] % S.leaddim == 0).all()
            gidx = np.where(whole[:, :, None],
                            rowd[:, :nf, None] + np.arange(csub), gidx)

        # every point is covered exactly once (periodic mesh: no
        # boundaries; the padding elements of the last block read their
        # own entries), sources are the views of the side beta selects
        live = np.arange(nb*csub).reshape(-1, csub) < 12
        live = live[:, None, :].repeat(nf, 1)
        assert (gidx[live] >= 0).all() and (gidx[~live] == -1).all()
        assert sorted(gidx[gidx >= 0]) == sorted(np.concatenate([src, src]))


def test_b200_unfused_plan_keeps_reference_decomposition(built):
    be, plan = _dry_plan('tgv', 2, {'fusion': 0}, order=2)
    kinds = [k for g in plan for w, k in g]

    assert kinds.count('mul') == 8          # disu, 2 grad, 3 fpts, 2 div
    assert 'copy' in kinds and 'gradflux' not in kinds


def test_b200_partitioned_plan_exchanges_once_per_graph(built):
    be, plan = _dry_plan('tgv', (4, 2, 2), nparts=2, order=2)
    xch = [[k for w, k in g if w == 'xchg'] for g in plan]

    # scal_fpts in the first graph, vect_fpts in the second, none after
    assert [len(x) for x in xch] == [1, 1, 0]
    assert sorted(xch[0][0]) == ['recv', 'send']

    # copy_fpts is elided here too: interior points are stored by intconu
    # (both sides), partition-boundary points by mpiconu in the next graph
    kinds = [k for g in plan for w, k in g if w == 'kernel']
    assert 'copy' not in kinds and 'mpiconu' in kinds


def test_b200_has_no_cpu_path(built):
    from pyfr_b200.backend import B200Backend
    from pyfr_b200.lib import B200NoDevice

    cfg, box = cases.make('tgv', 2, order=2)
    be = B200Backend(cfg, dry=True)
    s = get_system(be, box.local_mesh(), cfg, 2)

    with pytest.raises(B200NoDevice):
        s.rhs(0.0, 0, 1)


@pytest.mark.parametrize('beta,dead', [(0.5, True), (-0.5, True),
                                       (0.0, False)])
def test_b200_dead_gradient_rows(built, beta, dead):
    """With a one-sided LDG flux only one side of every interface reads
    the flux-point gradients; the fused kernel is then specialised not to
    compute or store the other half (and says so in its traffic count)."""
    from pyfr_b200.backend import B200Backend

    cfg, box = cases.make('tgv', 3, order=2, beta=beta)
    be = B200Backend(cfg, dry=True)
    s = get_system(be, box.local_mesh(), cfg, 2)
    gf, = [k for g in s.rhs_graphs(0, 1) for w, k in g.plan
           if w == 'kernel' and k.kind == 'gradflux']

    assert gf.info['dead_rows'] == dead

    nu, nf, LD, nb = 27, 54, 5*be.csubsz, -(-27 // be.csubsz)
    full = (2*nu + nf + 3*nf)*LD*nb*8
    # (folded common solution: + one 32-bit index per flux point)
    idx = (nf*be.csubsz + nf + 1)*nb*4 if gf.info['gather'] else 0
    assert gf.info['gather'] == dead
    assert gf.traffic == (full - 3*(nf // 2)*LD*nb*8 if dead else full) + idx


@pytest.mark.parametrize('rs', ['rusanov', 'hllc'])
@pytest.mark.parametrize('ndims', [2, 3])
def test_riemann_solver_consistency_and_symmetry(rs, ndims):
    """f(u, u, n) = F(u).n and f(l, r, n) = -f(r, l, -n)."""
    from oracle import physics as ph

    rng = np.random.default_rng(5)
    nv, m = ndims + 2, 64
    c = {'gamma': 1.4}

    def state():
        rho = rng.uniform(0.5, 2.0, m)
        vel = rng.uniform(-1.0, 1.0, (ndims, m))
        p = rng.uniform(0.5, 3.0, m)
        E = p/(c['gamma'] - 1) + 0.5*rho*(vel**2).sum(axis=0)
        return [rho, *(rho*v for v in vel), E]

    n = rng.standard_normal((ndims, m))
    n = list(n/np.sqrt((n**2).sum(axis=0)))
    ul, ur = state(), state()

    f, _, _ = ph.inviscid_flux(ul, ndims, nv, c)
    fn = ph.rsolvers[rs](ul, ul, n, ndims, nv, c)
    for i in range(nv):
        exact = sum(n[j]*f[j][i] for j in range(ndims))
        assert np.abs(fn[i] - exact).max() < 1e-12

    a = ph.rsolvers[rs](ul, ur, n, ndims, nv, c)
    b = ph.rsolvers[rs](ur, ul, [-x for x in n], ndims, nv, c)
    for i in range(nv):
        assert np.abs(a[i] + b[i]).max() < 1e-12


def test_affine_regions_are_detected(built):
    from pyfr_b200.backend import B200Backend

    def modes(**kw):
        cfg, box = cases.make('tgv', 3, order=2, **kw)
        s = get_system(B200Backend(cfg, dry=True), box.local_mesh(), cfg, 2)
        return [k.info['affine'] for g in s.rhs_graphs(0, 1)
                for w, k in g.plan if w == 'kernel' and k.kind == 'gradflux']

    assert modes() == [True]
    assert modes(warp=0.1) == [False]
    assert modes(curved=0.5) == [False, True]      # curved region, linear


def test_b200_boundary_plan(built):
    """Wall-bounded Navier-Stokes box: boundary kernels join the graphs
    where the reference puts them and copy_fpts stays elided (boundary
    points are stored by bcconu)."""
    from pyfr_b200.backend import B200Backend

    cfg, box, _ = cases.box_case('navier-stokes', (3, 3, 2),
                                 {'ylo': 'no-slp-adia-wall',
                                  'yhi': 'char-riem-inv'}, order=2)
    s = get_system(B200Backend(cfg, dry=True), box.local_mesh(), cfg, 2)
    graphs = s.rhs_graphs(0, 1)
    kinds = [[getattr(k, 'kind', None) for w, k in g.plan if w == 'kernel']
             for g in graphs]

    # (the interior intconu is folded into gradflux, which takes the
    # boundary points' common values from where bcconu stored them)
    assert kinds == [['mul', 'bcconu', 'bcconu'],
                     ['gradflux', None, 'bccflux', 'bccflux'],
                     ['mul+negdivconf']]

    gf = graphs[1].plan[0][1]
    gidx = gf.info['gidx'].get().reshape(3, 54, 8)
    nbc = 2*3*2*9                      # faces on ylo + yhi, points per face
    npad = 6*54                        # 18 elements in three blocks of 8
    assert (gidx == -1).sum() == nbc + npad


def test_unknown_boundary_type_is_refused():
    cfg, box, txt = cases.box_case('navier-stokes', (2, 2, 2),
                                   {'ylo': 'no-slp-adia-wall',
                                    'yhi': 'no-slp-adia-wall'}, order=1)
    cfg.set('soln-bcs-ylo', 'type', 'char-riem-inv-mass-flow')

    with pytest.raises(NotImplementedError, match='DESIGN.md'):
        get_system(OracleBackend(cfg), box.local_mesh(), cfg, 2)


def test_b200_euler_fused_plan(built):
    """With euler-fusion the Euler RHS is three launches: interpolation,
    interface flux, fused element kernel."""
    be, plan = _dry_plan('vortex', 6, {'euler-fusion': 1}, order=3)
    assert [[k for w, k in g] for g in plan] == [['mul', None], ['fluxdiv']]

    be, plan = _dry_plan('vortex', 6, {'euler-fusion': 0}, order=3)
    assert [k for w, k in plan[1]] == ['tflux', 'mul', 'mul+negdivconf'] or \
        [k for w, k in plan[1]] == ['tflux', 'mul', 'mul', 'negdivconf']


def test_irregular_partition_reproduces_single_partition():
    """Three irregular partitions made by the reference's partitioner
    (fixture conn_hex_periodic_3parts: uneven neighbour sets, several
    faces per neighbour pair) reproduce the single-partition RHS."""
    import os

    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                'golden', 'conn_hex_periodic_3parts.npz'))
    vparts, n = gold['vparts'], (5, 4, 3)
    kw = dict(order=2, beta=0.0, warp=0.1, rsolver='hllc')

    _, ref = oracle_rhs('tgv', n, **kw)
    _, out = oracle_rhs('tgv', n, vparts=vparts, nparts=3, **kw)

    _, box = cases.make('tgv', n, **kw)
    for r, o in enumerate(out):
        gidx = box.local_mesh(vparts, r).eidxs['hex']
        assert rel_err(o, ref[0][..., gidx]) < 5e-13


@pytest.mark.parametrize('case,n,kw', [
    ('tgv', (4, 3, 3), dict(order=2, warp=0.1)),
    ('tgv', (3, 3, 3), dict(order=3, beta=-0.5)),
    ('tgv', (3, 3, 3), dict(order=2, beta=0.0, rsolver='hllc')),
], ids=str)
def test_address_order_of_interface_points(case, n, kw):
    """inters-order = address: the points of the interior interfaces follow
    the true left-hand (for beta = -1/2: right-hand) address, i.e. the
    lanes of a block that share a flux-point row are consecutive; the
    reference's order (the default) keeps its own key, under which blocks
    interleave.  Either way the views hold the same set of points."""
    from pyfr_b200.backend import B200Backend

    maps = {}
    for order in ('reference', 'address'):
        cfg, box = cases.make(case, n, **kw)
        cfg.set('backend-b200', 'inters-order', order)
        cfg.set('backend-b200', 'conu-fold', 0)
        be = B200Backend(cfg, dry=True)
        sysm = get_system(be, box.local_mesh(), cfg, 2)
        conu, = [k for g in sysm.rhs_graphs(0, 1) for w, k in g.plan
                 if w == 'kernel' and getattr(k, 'kind', None) == 'intconu']
        side = 'urin' if kw.get('beta') == -0.5 else 'ulin'
        maps[order] = tuple(
            getattr(conu.info[s], 'view', conu.info[s]).mapping.get()[0]
            .astype(np.int64) for s in (side, 'ulin', 'urin')
        )

    ref, adr = maps['reference'], maps['address']
    assert np.all(np.diff(adr[0]) > 0)
    assert not np.all(np.diff(ref[0]) > 0) or len(ref[0]) < 64

    # Same interfaces: the (left, right) address pairs agree as sets
    pr = set(zip(ref[1].tolist(), ref[2].tolist()))
    pa = set(zip(adr[1].tolist(), adr[2].tolist()))
    assert pr == pa and len(pr) == len(ref[1])

    with pytest.raises(ValueError):
        cfg, box = cases.make(case, n, **kw)
        cfg.set('backend-b200', 'inters-order', 'random')
        B200Backend(cfg, dry=True)


# -- round 2: kernel-private point order, batched launches, generators ---------
def _dry_system(case, n, opts={}, vparts=None, rank=0, nparts=1, **kw):
    from pyfr_b200.backend import B200Backend

    cfg, box = cases.make(case, n, **kw)
    for k, v in opts.items():
        cfg.set('backend-b200', k, v)
    be = B200Backend(cfg, dry=True)
    comm = type('C', (), dict(rank=rank, size=nparts))()
    if vparts is not None and callable(vparts):
        vparts = vparts(box)
    return be, get_system(be, box.local_mesh(vparts, rank), cfg, 2, comm=comm)


def _plan_kernels(sysm):
    return [k for g in sysm.rhs_graphs(0, 1) for w, k in g.plan
            if w == 'kernel']


def test_interface_kernels_order_their_points_privately():
    """``kernel-order = address`` (the default): the interface kernels visit
    their points sorted by left-hand address through *copies* of the index
    arrays; the views the host built are the same objects with the same
    contents as under ``kernel-order = host``, so the reference's host code
    sees no difference."""
    maps = {}
    for order in ('host', 'address'):
        be, sysm = _dry_system('tgv', (4, 3, 3), {'kernel-order': order,
                                                  'conu-fold': 0},
                               order=2, warp=0.1)
        conu, = [k for k in _plan_kernels(sysm) if k.kind == 'intconu']
        v = conu.info['ulin']
        host = getattr(v, 'orig', getattr(v, 'view', v))
        maps[order] = (host.mapping.get()[0].copy(),
                       v.mapping.get()[0].copy(),
                       conu.info['urin'].mapping.get()[0].copy())

    (h0, k0, r0), (h1, k1, r1) = maps['host'], maps['address']
    assert np.array_equal(h0, h1)                    # host views untouched
    assert np.array_equal(h0, k0)                    # 'host': used as is
    assert not np.all(np.diff(h0) > 0)
    assert np.all(np.diff(k1) > 0)                   # sorted for the kernel
    # the same interfaces: (left, right) pairs agree as sets
    assert set(zip(k0.tolist(), r0.tolist())) == \
        set(zip(k1.tolist(), r1.tolist()))


def test_per_neighbour_kernels_are_batched():
    """Eight bricks: three neighbours per rank; ``pack``, ``mpiconu`` and
    ``mpicflux`` go out as at most two launches per kind (one per LDG
    orientation) instead of one per neighbour."""
    brick = lambda box: box.brick_partition((2, 2, 2))
    kinds = {}
    for batch in (0, 1):
        be, sysm = _dry_system('tgv', (4, 4, 4), {'batch-launches': batch},
                               vparts=brick, rank=3, nparts=8, order=2)
        ks = _plan_kernels(sysm)
        kinds[batch] = [k.kind or k.fn.name for k in ks]
        if batch:
            b = [k for k in ks if k.info and 'batched' in k.info]
            assert b and all(k.grid[1] == len(k.info['batched']) for k in b)
            npack = sum(len(k.info['batched']) for k in b
                        if k.kind == 'pack')

    # (solution: three neighbours; gradients: only to the neighbours this
    # rank is the sending side of)
    assert 4 <= kinds[0].count('pack') == npack <= 6
    assert kinds[0].count('mpicflux') == 3
    assert kinds[1].count('pack') == 2 and kinds[1].count('mpicflux') <= 2
    assert kinds[1].count('mpiconu') <= 2 < kinds[0].count('mpiconu')


def test_tensor_structure_is_verified_not_assumed():
    """``tp_structure`` accepts the operators of hexes and quads (and
    reproduces them from its 1-D factors) and declines everything else."""
    from pyfr_b200.host.shapes import shape_map
    from pyfr_b200.kernels.fused import NotFusable
    from pyfr_b200.kernels.tensor import tp_structure

    def ops(etype, order):
        cfg, _ = cases.make('tgv' if etype in ('hex', 'pri', 'tet', 'pyr')
                            else 'vortex', 2, order=order)
        if etype not in ('hex', 'quad'):
            cfg, _, _ = cases.mixed_case(
                'hex+pri+pyr+tet' if etype != 'tri' else 'quad+tri',
                (4, 4, 4) if etype != 'tri' else (4, 4), order=order)
        nverts = {'hex': 8, 'quad': 4, 'pri': 6, 'tet': 4, 'pyr': 5,
                  'tri': 3}[etype]
        b = shape_map[etype](nverts, cfg)
        return dict(A1=b.opmat('M4 - M6*M0'), M6=b.opmat('M6'),
                    M0=b.opmat('M0'), A5=b.opmat('M1 - M3*M2')), b.ndims

    for etype, order in (('hex', 1), ('hex', 2), ('hex', 4), ('quad', 3)):
        o, nd = ops(etype, order)
        st = tp_structure(o, nd)
        assert st['n1'] == order + 1 and st['nlines'] == (order + 1)**(nd - 1)
        assert all(s == (order + 1)**d for d, s in enumerate(st['stride']))

    for etype in ('pri', 'tet', 'pyr', 'tri'):
        o, nd = ops(etype, 3 if etype != 'pri' else 2)
        with pytest.raises(NotFusable):
            tp_structure(o, nd)

    # a perturbed hex operator (one entry outside the line structure)
    o, nd = ops('hex', 2)
    o['A5'] = o['A5'].copy()
    o['A5'][0, -1] += 1e-6
    with pytest.raises(NotFusable):
        tp_structure(o, nd)


def test_soa_width_follows_the_order():
    from pyfr_b200.backend import B200Backend

    want = {(4, 'double'): 8, (5, 'double'): 4, (6, 'single'): 4,
            (4, 'single'): 16, (2, 'double'): 8}
    for (order, prec), soa in want.items():
        cfg, _ = cases.make('tgv', 2, order=order, precision=prec)
        assert B200Backend(cfg, dry=True).soasz == soa

    cfg, _ = cases.make('tgv', 2, order=6, precision='single')
    cfg.set('backend-b200', 'n-soa', 16)
    assert B200Backend(cfg, dry=True).soasz == 16


def test_dense_kernel_takes_the_dense_operators_only():
    from pyfr_b200.kernels.dense import is_dense

    cfg, box, _ = cases.mixed_case('hex+pri+pyr+tet', (4, 4, 4), order=3)
    from pyfr_b200.backend import B200Backend

    be = B200Backend(cfg, dry=True)
    sysm = get_system(be, box.local_mesh(), cfg, 2)
    seen = {}
    for k in _plan_kernels(sysm):
        m = k.misc[0] if k.misc and isinstance(k.misc[0], dict) else {}
        if 'M' in m:
            seen[m['M'], m['K']] = bool(m.get('dense'))

    # pyramid / tet operators are dense, the hexahedron's are not
    assert seen[90, 56] and seen[30, 90] and seen[60, 20]
    assert not seen[96, 64] and not seen[64, 96]
    assert not is_dense(np.eye(200), 40, 8)
