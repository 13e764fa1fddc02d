"""The reference, executed: its own host code (``pyfr.solvers``) on its own
pointwise kernels -- every ``.mako`` kernel rendered by oracle/minimako.py,
completed by the reference's OpenMP kernel generator (argument
dereferencing of stacked matrices, views, broadcasts, 'mpi' arrays), built
with gcc and called through the reference's argument marshalling
(oracle/refkernels.py) -- against the same host code on the restated
NumPy oracle.  Only the matrix products, packing and register arithmetic
are shared (NumPy in place of libxsmm).  Needs /root/reference."""

import os
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import refharness as rh
from oracle.npbackend import LocalComm, make_backend
from pyfr_b200 import cases

pytestmark = pytest.mark.skipif(not rh.available(),
                                reason='needs /root/reference')

LAYOUT = '\n[backend-oracle]\nblocks = 1\nsoasz = 4\ncsubsz = 8\n'


def _systems(txt, meshes, mk, nregs=2, **kw):
    rh.install_stubs()
    import pyfr.backends.base as rbase
    from pyfr.inifile import Inifile
    from pyfr.solvers.euler import EulerSystem
    from pyfr.solvers.navstokes import NavierStokesSystem

    world = LocalComm(0, len(meshes))
    out = []
    for r, mesh in enumerate(meshes):
        rh.set_rank(world.peer(r))
        cfg = Inifile(txt + LAYOUT)
        be = mk(rbase)(cfg)
        cls = {'euler': EulerSystem, 'navier-stokes': NavierStokesSystem}[
            cfg.get('solver', 'system')]
        regs = [SimpleNamespace(rhs=True, dynamic=False, n=nregs,
                                extent=None)]
        s = cls(be, rh.ref_mesh(mesh), None, regs, cfg, None, **kw)
        s.commit()
        out.append(s)
    return out, world


def _rhs(systems, world, t=0.0):
    graphs = [s._rhs_graphs(0, 1) for s in systems]
    for s in systems:
        s._prepare_kernels(t, 0, 1)
    for stage in zip(*graphs):
        for g in stage:
            g.run()
        world.deliver()
    return [s.ele_scal_upts(1) for s in systems]


def _compare(txt, meshes, t=0.0, tol=1e-13):
    from oracle.refkernels import make_refkernel_backend

    res = [_rhs(*_systems(txt, meshes, mk), t=t)
           for mk in (make_backend, make_refkernel_backend)]

    for ro, rk in zip(*res):
        for a, b in zip(ro, rk):
            assert np.abs(a - b).max() <= tol*np.abs(a).max()
    return res


CASES = [
    ('vortex', 5, dict(order=3), (1, 1)),
    ('vortex', (6, 4), dict(order=3, rsolver='hllc'), (2, 1)),
    ('tgv', (3, 2, 2), dict(order=2, warp=0.1), (1, 1, 1)),
    ('tgv', (4, 2, 2), dict(order=2, warp=0.1, rsolver='hllc', beta=0.0),
     (2, 1, 1)),
    ('tgv', (3, 2, 2), dict(order=2, warp=0.1, beta=-0.5, curved=0.5,
                            visc_corr='sutherland'), (1, 1, 1)),
    ('tgv', (4, 2, 2), dict(order=3, warp=0.1, antialias='flux',
                            rsolver='hllc'), (2, 1, 1)),
    ('tgv', (3, 2, 2), dict(order=2, warp=0.1,
                            pts='gauss-legendre-lobatto'), (1, 1, 1)),
    # single precision (the generator suffixes every literal with f)
    ('tgv', (3, 2, 2), dict(order=2, warp=0.1, precision='single'),
     (1, 1, 1)),
    ('vortex', 5, dict(order=3, precision='single', rsolver='hllc'), (1, 1)),
]


@pytest.mark.parametrize('case,n,kw,parts', CASES, ids=str)
def test_oracle_equals_reference_kernels(case, n, kw, parts):
    kw2 = {k: v for k, v in kw.items() if k not in ('warp', 'curved')}
    txt = cases.tgv_cfg(**kw2) if case == 'tgv' else cases.vortex_cfg(**kw2)
    _, box = cases.make(case, n, **kw)

    nparts = int(np.prod(parts))
    vparts = box.brick_partition(parts) if nparts > 1 else None
    _compare(txt, [box.local_mesh(vparts, r) for r in range(nparts)],
             tol=1e-6 if kw.get('precision') == 'single' else 1e-13)


@pytest.mark.parametrize('system,n,bcs,kw', [
    ('navier-stokes', (3, 2, 2),
     {'xlo': 'sub-in-frv', 'xhi': 'sub-out-fp', 'ylo': 'no-slp-adia-wall',
      'yhi': 'char-riem-inv', 'zlo': 'slp-adia-wall',
      'zhi': 'no-slp-isot-wall'}, dict(order=2, warp=0.1, rsolver='hllc')),
    ('navier-stokes', (2, 3, 2),
     {'xlo': 'sub-in-ftpttang', 'xhi': 'sup-out-fn', 'ylo': 'sup-in-fa',
      'yhi': 'sub-out-fp'}, dict(order=2, beta=0.0)),
    ('euler', (5, 4),
     {'xlo': 'char-riem-inv', 'xhi': 'sup-out-fn', 'ylo': 'slp-adia-wall',
      'yhi': 'sup-in-fa'}, dict(order=3)),
], ids=str)
def test_boundary_conditions_equal_reference_kernels(system, n, bcs, kw):
    _, box, txt = cases.box_case(system, n, bcs, **kw)

    # a boundary value that depends on time and position
    if bcs['xlo'] == 'sub-in-frv':
        head, sect, tail = txt.partition('[soln-bcs-xlo]')
        tail = tail.replace('u = 0.2\n', 'u = 0.2 + 0.1*sin(3*t) + 0.05*y\n',
                            1)
        txt = head + sect + tail
        assert 'sin(3*t)' in txt

    _compare(txt, [box.local_mesh()], t=0.4)


def test_mixed_elements_equal_reference_kernels():
    _, box, txt = cases.mixed_case('hex+pri+pyr+tet', (4, 2, 2), order=2)
    _compare(txt, [box.local_mesh()])


def test_host_fixtures_are_what_the_reference_kernels_produce():
    """The committed host fixtures (tests/golden/host_*.npz) were recorded
    with the reference's host code on the *oracle* kernels; the reference's
    own kernels give the same numbers."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    import make_golden as mg

    for name in ('tgv_p2_rusanov', 'vortex_p3_rusanov'):
        case, n, kw, parts, beopts = mg.HOST_CASES[name]
        gold = np.load(os.path.join(os.path.dirname(__file__), 'golden',
                                    f'host_{name}.npz'))
        _, box = cases.make(case, n, **kw)

        from oracle.refkernels import make_refkernel_backend
        (rhs,), = [_rhs(*_systems(mg.cfg_text(case, kw, {}),
                                  [box.local_mesh()],
                                  make_refkernel_backend))]
        assert np.abs(rhs[0] - gold['r0_rhs']).max() <= \
            1e-13*np.abs(gold['r0_rhs']).max()


@pytest.mark.parametrize('name', ['vortex_p3_rk45_pi_l2',
                                  'tgv_p2_rk45_cfl_curved'])
def test_integrator_fixtures_hold_on_reference_kernels(name):
    """The reference's integrators (RK45 under the PI and CFL controllers)
    on the reference's own ``rkvdh2`` / ``wavespeed`` / RHS kernels retrace
    the committed histories (tests/golden/intg_*.npz, recorded on the
    oracle kernels)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    import make_golden as mg

    from oracle.refkernels import make_refkernel_backend

    rh.install_stubs()
    rh.set_rank(LocalComm(0, 1).peer(0))
    import pyfr.backends.base as rbase
    from pyfr.inifile import Inifile
    from pyfr.integrators import get_integrator
    from pyfr.solvers.euler import EulerSystem
    from pyfr.solvers.navstokes import NavierStokesSystem

    case, n, kw, opts, tlist = mg.INTG_CASES[name]
    _, box = cases.make(case, n, **kw)
    cfg = Inifile(mg.intg_cfg_text(name) + LAYOUT)
    be = make_refkernel_backend(rbase)(cfg)
    cls = EulerSystem if case == 'vortex' else NavierStokesSystem
    intg = get_integrator(be, cls, rh.ref_mesh(box.local_mesh()), None, cfg)

    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden',
                                f'intg_{name}.npz'))
    for i, t in enumerate(tlist):
        intg.advance_to(t)
        assert intg.tcurr == float(gold[f'tcurr_t{i}'])
        u = intg.soln[0]
        assert np.abs(u - gold[f'u_t{i}']).max() <= \
            1e-12*np.abs(gold[f'u_t{i}']).max()

    assert (intg.nacptsteps, intg.nrjctsteps) == tuple(gold['counts'][:2])
    assert intg.dt == pytest.approx(float(gold['dt_final']), rel=1e-9)


_direct = r'''
import os, sys
os.environ['PYFR_B200_BASE'] = 'pyfr.backends.base'
ROOT = %(root)r
sys.path[:0] = [ROOT, ROOT + '/tests', ROOT + '/tests/cudaemu']
from types import SimpleNamespace
import numpy as np
from oracle import refharness as rh
from oracle.npbackend import LocalComm
from oracle.refkernels import make_refkernel_backend
rh.install_stubs()
rh.set_rank(LocalComm(0, 1).peer(0))
import emu
import pyfr_b200.backend as bk, pyfr_b200.compiler as comp
bk.load_runtime = lambda device=0, dry=False: emu.EmuRuntime()
comp.KernelCompiler.cubin = lambda self, src, name: src.encode()
import pyfr.backends.base as rbase
from pyfr.inifile import Inifile
from pyfr.solvers.euler import EulerSystem
from pyfr.solvers.navstokes import NavierStokesSystem
from pyfr_b200 import cases
from pyfr_b200.backend import B200Backend

for case, n, kw in [('tgv', (3, 2, 2), dict(order=3, warp=0.1)),
                    ('tgv', (3, 2, 2), dict(order=2, warp=0.1, beta=0.0,
                                            rsolver='hllc')),
                    ('vortex', 5, dict(order=3))]:
    kw2 = {k: v for k, v in kw.items() if k != 'warp'}
    txt = (cases.tgv_cfg(**kw2) if case == 'tgv' else cases.vortex_cfg(**kw2))
    txt += '\n[backend-oracle]\nblocks = 1\nsoasz = 4\ncsubsz = 8\n'
    _, box = cases.make(case, n, **kw)
    mesh = rh.ref_mesh(box.local_mesh())
    cls = NavierStokesSystem if case == 'tgv' else EulerSystem

    outs = []
    for mk in (lambda cfg: make_refkernel_backend(rbase)(cfg), B200Backend):
        cfg = Inifile(txt)
        regs = [SimpleNamespace(rhs=True, dynamic=False, n=2, extent=None)]
        s = cls(mk(cfg), mesh, None, regs, cfg, None)
        s.commit()
        s.rhs(0.0, 0, 1)
        outs.append(s.ele_scal_upts(1)[0])

    print('RESULT', case, np.abs(outs[1] - outs[0]).max()/np.abs(outs[0]).max())
'''


def test_b200_backend_against_reference_kernels_directly():
    """Same unmodified reference host code, two backends: the reference's
    own generated kernels, and the B200 backend (its generated CUDA run on
    the CPU execution model, fused launches and all)."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, '-c', _direct % {'root': root}],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2500:]

    rows = [l.split() for l in res.stdout.splitlines()
            if l.startswith('RESULT')]
    assert len(rows) == 3 and all(float(r[2]) < 1e-12 for r in rows), rows
