"""GPU parity: the B200 backend against the NumPy oracle on identical
inputs, through the backend API (every launch goes through the C ABI)."""

import numpy as np
import pytest

from pyfr_b200 import cases
from pyfr_b200.host.system import get_system

from util import assert_parity, oracle_rhs, rel_err, rhs_magnitude

pytestmark = pytest.mark.gpu

# fp64 per-point RHS tolerance stated by BASELINE.json's north_star,
# normalised by the field's infinity norm
TOL64 = 1e-12


# Backend variants: default (64-byte SoA, fused kernels), the wide layout
# (fused element kernel does not fit -> individual kernels) and fusion off
VARIANTS = {'default': {}, 'soa16': {'n-soa': 16}, 'nofuse': {'fusion': 0},
            'soa4': {'n-soa': 4}}


def b200_rhs(case, n, opts={}, **kw):
    from pyfr_b200.backend import B200Backend

    cfg, box = cases.make(case, n, **kw)
    for k, v in opts.items():
        cfg.set('backend-b200', k, v)
    be = B200Backend(cfg)
    sysm = get_system(be, box.local_mesh(), cfg, 2)
    sysm.rhs(0.0, 0, 1)
    be.wait()

    return sysm, sysm.ele_scal_upts(1)[0]


@pytest.mark.parametrize('kw', [
    dict(order=2), dict(order=2, rsolver='hllc'), dict(order=2, beta=0.0),
    dict(order=2, beta=-0.5), dict(order=3, rsolver='hllc'), dict(order=4),
    dict(order=4, rsolver='hllc'), dict(order=3, curved=0.5),
    dict(order=2, curved=1.0, rsolver='hllc')
], ids=str)
@pytest.mark.parametrize('variant', list(VARIANTS))
def test_tgv_rhs_matches_oracle(built, kw, variant):
    n = (5, 4, 3)
    _, ref = oracle_rhs('tgv', n, warp=0.1, **kw)
    esys, ext = oracle_rhs('tgv', n, warp=0.1, extended=True, **kw)
    sysm, out = b200_rhs('tgv', n, VARIANTS[variant], warp=0.1, **kw)

    assert out.shape == ref[0].shape
    assert_parity(out, ref[0], ext[0], TOL64, mag=rhs_magnitude(esys[0])[0])

    kinds = [getattr(k, 'kind', None) for g in sysm.rhs_graphs(0, 1)
             for w, k in g.plan if w == 'kernel']
    if variant in ('default', 'soa4'):
        assert 'gradflux' in kinds and 'mul+negdivconf' in kinds
        assert 'copy' not in kinds
    elif variant == 'nofuse':
        assert 'gradflux' not in kinds
        assert ('copy' in kinds) == (abs(kw.get('beta', 0.5)) == 0.5)
    elif kw['order'] == 4:
        # 16-wide fp64 blocks at p=4 do not fit the fused kernel's smem
        assert 'gradflux' not in kinds


@pytest.mark.parametrize('kw', [dict(order=3), dict(order=3, rsolver='hllc')],
                         ids=str)
def test_vortex_rhs_matches_oracle(built, kw):
    _, ref = oracle_rhs('vortex', 12, **kw)
    esys, ext = oracle_rhs('vortex', 12, extended=True, **kw)
    _, out = b200_rhs('vortex', 12, {}, **kw)

    assert_parity(out, ref[0], ext[0], TOL64, mag=rhs_magnitude(esys[0])[0])


def test_matrix_roundtrip_and_layout(built):
    from pyfr_b200.backend import B200Backend

    cfg, _ = cases.make('tgv', 2)
    be = B200Backend(cfg)
    rng = np.random.default_rng(0)

    for shape in [(7, 5, 37), (3, 4, 5, 19), (6, 50)]:
        a = rng.standard_normal(shape)
        m = be.matrix(shape, a, tags={'align'})
        assert np.array_equal(m.get(), a)

        b = rng.standard_normal(shape)
        m.set(b)
        assert np.array_equal(m.get(), b)


def test_free_stream_preserved(built):
    extra = ''
    from pyfr_b200.backend import B200Backend
    from pyfr_b200.host.config import Config

    txt = cases.tgv_cfg(order=3)
    txt = txt[:txt.index('[soln-ics]')] + (
        '[soln-ics]\nrho = 1.2\nu = 0.3\nv = -0.2\nw = 0.1\np = 2.5\n'
    )
    cfg = Config(txt)
    be = B200Backend(cfg)
    box = cases.tgv_mesh((4, 3, 3), warp=0.15)
    sysm = get_system(be, box.local_mesh(), cfg, 2)
    sysm.rhs(0.0, 0, 1)

    assert np.abs(sysm.ele_scal_upts(1)[0]).max() < 1e-11


@pytest.mark.parametrize('kw', [dict(order=2), dict(order=4),
                                dict(order=3, rsolver='hllc', beta=0.0),
                                # BASELINE configs[4] proxy: p = 6 (SoA width
                                # 4, sum-factorised fused kernel)
                                dict(order=6), dict(order=6, rsolver='hllc')],
                         ids=str)
def test_tgv_rhs_fp32(built, kw):
    """Single precision (the layout doubles the SoA width to 16): held to
    the fp32 oracle's own distance from the fp64 oracle -- at M = 0.1 the
    RHS is a cancellation of O(1/M^2) terms, so fp32 round-off alone is
    1e-4..1e-3 of the field maximum."""
    n = (4, 3, 3)
    _, r64 = oracle_rhs('tgv', n, warp=0.1, **kw)
    _, r32 = oracle_rhs('tgv', n, warp=0.1, precision='single', **kw)
    _, out = b200_rhs('tgv', n, {}, warp=0.1, precision='single', **kw)

    assert out.dtype == np.float32
    floor = rel_err(r32[0].astype(float), r64[0])
    err = rel_err(out.astype(float), r64[0])

    from util import PARITY_LOG
    PARITY_LOG.append(dict(test=f'fp32 {kw}', err=float(err),
                           floor=float(floor), ratio=None,
                           ratio_oracle=None))
    assert err <= max(4*floor, 1e-5), (err, floor)


@pytest.mark.parametrize('kw', [dict(order=2), dict(order=4),
                                dict(order=3, rsolver='hllc', beta=0.0),
                                dict(order=4, precision='single')],
                         ids=str)
@pytest.mark.parametrize('opts', [{}, {'gradflux-threads': 640}], ids=str)
def test_tgv_rhs_affine_mesh(built, kw, opts):
    """Uniform (parallelepiped) elements take the constant-Jacobian fast
    path of the fused kernel: the metric terms live in registers."""
    n = (4, 3, 3)
    _, ref = oracle_rhs('tgv', n, **kw)
    sysm, out = b200_rhs('tgv', n, opts, **kw)

    gf, = [k for g in sysm.rhs_graphs(0, 1) for w, k in g.plan
           if w == 'kernel' and k.kind == 'gradflux']
    assert gf.info['affine']

    if kw.get('precision') == 'single':
        _, r64 = oracle_rhs('tgv', n, **{**kw, 'precision': 'double'})
        floor = rel_err(ref[0].astype(float), r64[0])
        assert rel_err(out.astype(float), r64[0]) <= max(4*floor, 1e-5)
    else:
        esys, ext = oracle_rhs('tgv', n, extended=True, **kw)
        assert_parity(out, ref[0], ext[0], TOL64,
                      mag=rhs_magnitude(esys[0])[0])


BC_CASES = [
    ('navier-stokes', (4, 3, 3), {'ylo': 'no-slp-adia-wall',
                                  'yhi': 'char-riem-inv'},
     dict(order=3, warp=0.1)),
    ('navier-stokes', (3, 4, 3), {'xlo': 'sub-in-frv', 'xhi': 'sub-out-fp',
                                  'zlo': 'slp-adia-wall',
                                  'zhi': 'no-slp-isot-wall'},
     dict(order=2, rsolver='hllc')),
    ('navier-stokes', (3, 3, 4), {'xlo': 'sup-in-fa', 'xhi': 'sup-out-fn',
                                  'ylo': 'no-slp-adia-wall',
                                  'yhi': 'no-slp-adia-wall'},
     dict(order=2, beta=0.0, warp=0.05)),
    ('euler', (9, 8), {'xlo': 'char-riem-inv', 'xhi': 'sup-out-fn',
                       'ylo': 'slp-adia-wall', 'yhi': 'sup-in-fa'},
     dict(order=3)),
    ('euler', (8, 8), {'ylo': 'slp-adia-wall', 'yhi': 'slp-adia-wall'},
     dict(order=2, rsolver='hllc')),
]


@pytest.mark.parametrize('system,n,bcs,kw', BC_CASES, ids=str)
def test_boundary_conditions_match_oracle(built, system, n, bcs, kw):
    """bcconu / bccflux for every boundary type on the path, wall-bounded
    and open boxes, against the oracle driven by the same host code."""
    from pyfr_b200.backend import B200Backend
    from util import OracleBackend

    def run(cls, extended=False):
        cfg, box, _ = cases.box_case(system, n, bcs, **kw)
        cfg.set('backend-oracle', 'extended-mul', extended)
        sysm = get_system(cls(cfg), box.local_mesh(), cfg, 2)
        sysm.rhs(0.25, 0, 1)
        if hasattr(sysm.backend, 'wait'):
            sysm.backend.wait()
        return sysm, sysm.ele_scal_upts(1)[0]

    _, ref = run(OracleBackend)
    esys, ext = run(OracleBackend, extended=True)
    sysm, out = run(B200Backend)

    # (no point-wise running-error criterion here: its magnitude field is
    # built from the interior trace alone, while a boundary flux is summed
    # from the *ghost* state's terms -- prescribed inflow values, pow() of
    # the characteristic boundary -- which can exceed it many times over)
    assert_parity(out, ref, ext, TOL64)

    kinds = [getattr(k, 'kind', None) for g in sysm.rhs_graphs(0, 1)
             for w, k in g.plan if w == 'kernel']
    assert 'bccflux' in kinds and 'copy' not in kinds


def test_full_size_conservation_and_symmetry(built):
    """BASELINE.json config #2 at full size (64^3 hexes, p=4, fp64: 164 M
    DoF), checked through size-independent properties that need no
    reference evaluation: the RHS integrates to zero over the periodic box
    for every conserved variable (discrete conservation), and the RHS of
    the mass equation inherits the TGV initial condition's symmetry
    (rho = 1, div u = 0 -> d rho/dt is pure round-off relative to the
    momentum terms)."""
    from util import conservation_defect

    cfg, box = cases.make('tgv', 64, order=4)
    from pyfr_b200.backend import B200Backend

    be = B200Backend(cfg)
    mesh = box.local_mesh()
    sysm = get_system(be, mesh, cfg, 2)
    sysm.rhs(0.0, 0, 1)
    be.wait()
    rhs = sysm.ele_scal_upts(1)[0]

    assert rhs.shape == (125, 5, 64**3)
    assert np.isfinite(rhs).all()

    tot, mag = conservation_defect(cfg, mesh, rhs)
    assert np.all(np.abs(tot) <= 1e-11*mag.max()), (tot, mag)

    # x-momentum and y-momentum magnitudes agree by the IC's symmetry
    assert abs(mag[1]/mag[2] - 1) < 1e-10
    # no mass sources at t = 0 beyond discretisation error of div u
    assert np.abs(rhs[:, 0]).max() < 1e-4*np.abs(rhs[:, 1]).max()


def test_full_size_rhs_matches_c_oracle(built):
    """BASELINE.json config #2 at full size (64^3 hexes, p=4, fp64, 164 M
    DoF): one RHS on the device against the C restatement of the path
    (oracle/crhs, the build *without* -ffast-math) on identical inputs.

    What can be asserted at this size.  The RHS of the TGV initial
    condition at M = 0.1 is O(1) while the terms it is summed from --
    interface terms ``lambda E`` times the correction operator times the
    metric -- are O(n/M^2): the distance of *any* fp64 evaluation from the
    exact value grows in proportion to the mesh size n (NumPy oracle against
    its own extended-precision evaluation, relative to the field maximum:
    6e-12 at 4^3, 1.7e-11 at 6^3, 6.1e-11 at 16^3 => ~2.5e-10 at 64^3), so
    "1e-12 of the field maximum" is not a property an fp64 backend can have
    here.  The scale-free statement is the point-wise one: every point lies
    within a modest multiple of ``eps`` times the magnitude of its own terms
    (tests/util.py: rhs_magnitude; both oracles sit at 3-4.4 for every n).
    Two fp64 evaluations may therefore differ by up to twice that.  At
    this size a second effect shows: the metric terms are differences of
    vertex coordinates, exact only to ``eps |x|/h`` (32 eps at the corners
    of this box), and every term of the RHS is linear in them.  Which
    rounding one gets depends on how the Jacobian is formed -- the C oracle
    and the device kernels form it differently -- so against the C oracle
    the criterion allows for it (tests/util.py: geometry_conditioning),
    while two *independent device kernels* (the sum-factorised affine
    kernel and the table-driven general-geometry kernel under the host's
    interface order), which form the Jacobian alike, must meet the plain
    limit ``RUNNING_ERROR_C``.  Measured (r02g): 58.9 eps against the C
    oracle without the allowance, 58.2 for the general kernel against it,
    5.3 between the two kernels.  Checked on every point of a quarter of
    the mesh (every fourth element block); the field-maximum error, the
    relative L2 error per variable and the bias of the differences are
    recorded."""
    import gc
    import os

    from oracle.cbackend import make_cbackend
    from pyfr_b200 import base
    from pyfr_b200.backend import B200Backend
    from util import (PARITY_LOG, RUNNING_ERROR_C, geometry_conditioning,
                      rhs_magnitude_from_state, running_error_ratio)

    n = 64
    cfg, box = cases.make('tgv', n, order=4)
    mesh = box.local_mesh()
    sysm = get_system(B200Backend(cfg), mesh, cfg, 2)
    u0 = sysm.ele_scal_upts(0)[0]
    sysm.rhs(0.0, 0, 1)
    sysm.backend.wait()
    out = sysm.ele_scal_upts(1)[0]
    del sysm
    gc.collect()

    CB = make_cbackend(base, fast=False, nthreads=os.cpu_count())
    cfg, box = cases.make('tgv', n, order=4)
    csys = get_system(CB(cfg), mesh, cfg, 2)
    csys.rhs(0.0, 0, 1)
    ref = csys.ele_scal_upts(1)[0]
    del csys
    gc.collect()

    assert out.shape == ref.shape == (125, 5, n**3)
    assert np.isfinite(out).all()

    # point-wise criterion on every fourth block of eight elements
    eidx = np.flatnonzero((np.arange(n**3)//8) % 4 == 0)
    mag = rhs_magnitude_from_state(cfg, mesh, u0, eidx)
    ratio_raw = running_error_ratio(out[..., eidx], ref[..., eidx], mag)
    geo = geometry_conditioning(mesh, eidx)
    ratio = running_error_ratio(out[..., eidx], ref[..., eidx], mag*geo)

    d = out - ref
    err = float(np.abs(d).max()/np.abs(ref).max())
    l2 = [float(np.linalg.norm(d[:, v])/np.linalg.norm(ref[:, v]))
          for v in range(1, 5)]
    bias = [float(abs(d[:, v].mean())/max(d[:, v].std(), 1e-300))
            for v in range(1, 5)]
    rec = dict(test='full-size 64^3 p=4 vs oracle/crhs',
               err=err, floor=float('nan'), ratio=ratio,
               ratio_oracle=None, l2_momentum_energy=l2,
               ratio_without_geometry_allowance=ratio_raw,
               bias=bias, npoints_checked=int(mag.size))
    PARITY_LOG.append(rec)

    # Where the largest point-wise deviation sits, its distribution, and a
    # third evaluation (the table-driven general-geometry kernel with the
    # host's interface order) to tell which of the two sides it belongs to
    q = np.abs(out[..., eidx] - ref[..., eidx])/(
        np.finfo(float).eps*np.maximum(mag, 1e-300))
    pmax = np.unravel_index(np.argmax(q), q.shape)
    rec['ratio_percentiles'] = {
        str(p): float(np.percentile(q, p)) for p in (50, 99, 99.99)}
    rec['worst_point'] = dict(upt=int(pmax[0]), var=int(pmax[1]),
                              ele=int(eidx[pmax[2]]))

    cfg, box = cases.make('tgv', n, order=4)
    for k, v in (('gradflux-tensor', 0), ('affine-fastpath', 0),
                 ('kernel-order', 'host')):
        cfg.set('backend-b200', k, v)
    sys2 = get_system(B200Backend(cfg), mesh, cfg, 2)
    sys2.rhs(0.0, 0, 1)
    sys2.backend.wait()
    out2 = sys2.ele_scal_upts(1)[0]
    del sys2
    gc.collect()
    rec['ratio_general_kernel_vs_crhs'] = running_error_ratio(
        out2[..., eidx], ref[..., eidx], mag)
    rec['ratio_vs_general_kernel'] = running_error_ratio(
        out[..., eidx], out2[..., eidx], mag)
    assert rec['ratio_vs_general_kernel'] <= RUNNING_ERROR_C

    assert ratio <= RUNNING_ERROR_C, ratio
    # (a loose global sanity bound on top: 64^3 sits at ~2.5e-10)
    assert err < 5e-9, err


@pytest.mark.parametrize('kw', [dict(order=3), dict(order=2, rsolver='hllc'),
                                dict(order=4, precision='single')], ids=str)
def test_vortex_rhs_fused_euler_kernel(built, kw):
    """The fused Euler element kernel (flux + divergence + correction +
    scaling in one launch): three launches per RHS."""
    _, ref = oracle_rhs('vortex', 12, **kw)
    sysm, out = b200_rhs('vortex', 12, {'euler-fusion': 1}, **kw)

    kinds = [getattr(k, 'kind', None) for g in sysm.rhs_graphs(0, 1)
             for w, k in g.plan if w == 'kernel']
    assert kinds.count('fluxdiv') == 1 and 'tflux' not in kinds

    if kw.get('precision') == 'single':
        _, r64 = oracle_rhs('vortex', 12, **{**kw, 'precision': 'double'})
        floor = rel_err(ref[0].astype(float), r64[0])
        assert rel_err(out.astype(float), r64[0]) <= max(4*floor, 1e-5)
    else:
        _, ext = oracle_rhs('vortex', 12, extended=True, **kw)
        assert_parity(out, ref[0], ext[0], TOL64)


def test_euler_boundaries_fused_kernel(built):
    from pyfr_b200.backend import B200Backend
    from util import OracleBackend

    system, n, bcs, kw = BC_CASES[3]
    outs = []
    for cls, ext in ((OracleBackend, False), (OracleBackend, True),
                     (B200Backend, False)):
        cfg, box, _ = cases.box_case(system, n, bcs, warp=0.1, **kw)
        cfg.set('backend-oracle', 'extended-mul', ext)
        cfg.set('backend-b200', 'euler-fusion', 1)
        sysm = get_system(cls(cfg), box.local_mesh(), cfg, 2)
        sysm.rhs(0.0, 0, 1)
        if hasattr(sysm.backend, 'wait'):
            sysm.backend.wait()
        outs.append(sysm.ele_scal_upts(1)[0])

    assert_parity(outs[2], outs[0], outs[1], TOL64)
