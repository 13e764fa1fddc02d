"""The backend as a drop-in under the REFERENCE's own classes.

Runs only where /root/reference exists (build container).  In a
subprocess with ``PYFR_B200_BASE=pyfr.backends.base`` the B200 backend
derives from the reference's ``BaseBackend``/``Matrix``/``View``/``Graph``
and is driven by the reference's unmodified ``NavierStokesSystem`` /
``EulerSystem`` (``pyfr/solvers/*``): every ``backend.kernel(...)`` call,
view, exchange registration and ``Graph.group`` hint then comes from the
reference's host code.  Without a GPU nothing can execute (dry runtime), so
the check is that set-up, kernel generation, fusion and graph commit all
succeed and yield the same launch plan as with this repository's host
mirror."""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_available():
    sys.path.insert(0, ROOT)
    try:
        from oracle import refharness
        return refharness.available()
    finally:
        sys.path.pop(0)

_script = r'''
import json, os, sys
os.environ['PYFR_B200_BASE'] = 'pyfr.backends.base'
sys.path.insert(0, %(root)r)
from types import SimpleNamespace
from oracle import refharness as rh
from oracle.npbackend import LocalComm
rh.install_stubs()
rh.set_rank(LocalComm(0, 1))
from pyfr.inifile import Inifile
from pyfr.backends.base import BaseBackend
from pyfr.solvers.euler import EulerSystem
from pyfr.solvers.navstokes import NavierStokesSystem
from pyfr.util import subclass_where
from pyfr_b200 import cases
from pyfr_b200.backend import B200Backend

assert subclass_where(BaseBackend, name='b200') is B200Backend

out = {}
for case, n, kw in [('tgv', 3, dict(order=4)), ('tgv', 3, dict(order=2, beta=0.0)),
                    ('vortex', 4, dict(order=3))]:
    txt = (cases.tgv_cfg(**kw) if case == 'tgv' else cases.vortex_cfg(**kw))
    cfg = Inifile(txt)
    _, box = cases.make(case, n, **kw)
    be = B200Backend(cfg, dry=True)
    regs = [SimpleNamespace(rhs=True, dynamic=False, n=2, extent=None)]
    cls = NavierStokesSystem if case == 'tgv' else EulerSystem
    s = cls(be, rh.ref_mesh(box.local_mesh()), None, regs, cfg, None)
    s.commit()
    out[f'{case}{kw}'] = [[getattr(k, 'kind', None) for w, k in g.plan if w == 'kernel']
                          for g in s._rhs_graphs(0, 1)]
print('RESULT ' + json.dumps(out))
'''


@pytest.mark.skipif(not _ref_available(),
                    reason='needs /root/reference')
def test_b200_backend_under_reference_host(built):
    res = subprocess.run([sys.executable, '-c', _script % {'root': ROOT}],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]

    line, = [l for l in res.stdout.splitlines() if l.startswith('RESULT ')]
    plans = json.loads(line[7:])

    ns4, ns2, eu = plans.values()
    # (intconu is folded into gradflux: the reference's g1 is re-planned
    # when its g2 is committed)
    assert ns4 == [['mul'], ['gradflux', None], ['mul+negdivconf']]
    # (beta = 0: the common solution is an average, intconu stays)
    assert ns2 == [['mul', 'intconu'], ['gradflux', None], ['mul+negdivconf']]
    assert [k for g in eu for k in g].count(None) >= 1      # intcflux
    assert eu[-1] == ['fluxdiv']


@pytest.mark.skipif(not _ref_available(),
                    reason='needs /root/reference')
def test_memory_info_under_reference_base(built):
    """memory_info() goes through whichever base class is in use."""
    code = r'''
import os, sys
os.environ['PYFR_B200_BASE'] = 'pyfr.backends.base'
sys.path.insert(0, %(root)r)
from oracle import refharness as rh
rh.install_stubs()
from pyfr.inifile import Inifile
from pyfr_b200 import cases
from pyfr_b200.backend import B200Backend
be = B200Backend(Inifile(cases.tgv_cfg(order=2)), dry=True)
be.matrix((4, 5, 8), tags={'align'})
be.commit()
mi = be.memory_info()
assert mi.current >= 4*5*8*8 and mi.peak >= mi.current, mi
print('OK')
'''
    res = subprocess.run([sys.executable, '-c', code % {'root': ROOT}],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and 'OK' in res.stdout, res.stderr[-1500:]


_exec_script = r'''
import os, sys
os.environ['PYFR_B200_BASE'] = 'pyfr.backends.base'
ROOT = %(root)r
sys.path[:0] = [ROOT, ROOT + '/tests', ROOT + '/tests/cudaemu']
from types import SimpleNamespace
import numpy as np
from oracle import refharness as rh
from oracle.npbackend import LocalComm
rh.install_stubs()
import emu
import pyfr_b200.backend as bk, pyfr_b200.compiler as comp
bk.load_runtime = lambda device=0, dry=False: emu.EmuRuntime()
comp.KernelCompiler.cubin = lambda self, src, name: src.encode()
from pyfr.inifile import Inifile
from pyfr.solvers.navstokes import NavierStokesSystem
from pyfr.solvers.euler import EulerSystem
from pyfr_b200 import cases
from pyfr_b200.backend import B200Backend

for case, n, kw in [('tgv', (3, 2, 2), dict(order=2, warp=0.1)),
                    ('vortex', 5, dict(order=3)),
                    ('tgv', (3, 2, 2), dict(order=2, warp=0.1,
                                            antialias='flux')),
                    ('vortex', 5, dict(order=3, antialias='flux')),
                    ('tgv', (3, 2, 2), dict(order=3, warp=0.1,
                                            visc_corr='sutherland',
                                            rsolver='hllc')),
                    ('tgv', (3, 2, 2), dict(order=2, warp=0.1,
                                            antialias='flux, surf-flux')),
                    ('vortex', 5, dict(order=3, antialias='surf-flux'))]:
    kw2 = {k: v for k, v in kw.items() if k != 'warp'}
    txt = (cases.tgv_cfg(**kw2) if case == 'tgv' else cases.vortex_cfg(**kw2))
    txt += '\n[backend-b200]\ngraphs = false\n'
    if 'surf-flux' in kw.get('antialias', ''):
        # quadrature degree of the surface (and volume) anti-aliasing rules
        txt = txt.replace('pts = gauss-legendre\n',
                          'pts = gauss-legendre\nquad-deg = 7\n')
    _, box = cases.make(case, n, **kw)
    mesh = box.local_mesh()
    world = LocalComm(0, 1)
    rh.set_rank(world.peer(0))

    be = B200Backend(Inifile(txt))
    regs = [SimpleNamespace(rhs=True, dynamic=False, n=2, extent=None)]
    cls = NavierStokesSystem if case == 'tgv' else EulerSystem
    s = cls(be, rh.ref_mesh(mesh), None, regs, Inifile(txt), None)
    s.commit()
    s.rhs(0.0, 0, 1)
    out = s.ele_scal_upts(1)[0]

    rs, rbe = rh.ref_system(txt, mesh, 2, world.peer(0))
    rs.rhs(0.0, 0, 1)
    ref = rs.ele_scal_upts(1)[0]
    print('RESULT', case, np.abs(out - ref).max()/np.abs(ref).max(),
          be.rt.nlaunch)
'''


@pytest.mark.skipif(not _ref_available(),
                    reason='needs /root/reference')
def test_reference_host_executes_on_b200_backend(built):
    """The complete drop-in, executed: the reference's unmodified systems
    drive B200Backend (derived from the reference's base classes), whose
    generated CUDA kernels run on the CPU execution model; the RHS equals
    the reference host + oracle backend result, in 4 (Navier-Stokes) and
    3 (Euler) launches."""
    res = subprocess.run([sys.executable, '-c',
                          _exec_script % {'root': ROOT}],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]

    rows = [l.split() for l in res.stdout.splitlines()
            if l.startswith('RESULT')]
    # + flux anti-aliasing (both systems) and Sutherland's law with HLLC
    # ... and surface-flux anti-aliasing (host mirror does not build those
    # operators; the reference's shapes do)
    assert [r[1] for r in rows] == ['tgv', 'vortex', 'tgv', 'vortex', 'tgv',
                                    'tgv', 'vortex']
    assert all(float(r[2]) < 1e-12 for r in rows)
    assert [int(r[3]) for r in rows][:2] == [4, 3]
    assert int(rows[4][3]) == 4


_intg_script = r'''
import os, sys
os.environ['PYFR_B200_BASE'] = 'pyfr.backends.base'
ROOT = %(root)r
sys.path[:0] = [ROOT, ROOT + '/tests', ROOT + '/tests/cudaemu']
import numpy as np
from oracle import refharness as rh
from oracle.npbackend import LocalComm, make_backend
rh.install_stubs()
rh.set_rank(LocalComm(0, 1).peer(0))
import emu
import pyfr_b200.backend as bk, pyfr_b200.compiler as comp
bk.load_runtime = lambda device=0, dry=False: emu.EmuRuntime()
comp.KernelCompiler.cubin = lambda self, src, name: src.encode()
import pyfr.backends.base as rbase
from pyfr.inifile import Inifile
from pyfr.integrators import get_integrator
from pyfr.solvers.euler import EulerSystem
from pyfr.solvers.navstokes import NavierStokesSystem
from pyfr_b200 import cases
from pyfr_b200.backend import B200Backend

INTG = {
    'pi': ('vortex', (4, 4), dict(order=3), 'scheme = rk45\ncontroller = pi\n'
           'dt = 0.05\natol = 1e-6\nrtol = 1e-6\n', 0.12),
    'cfl': ('tgv', (2, 2, 2), dict(order=2, warp=0.1), 'scheme = rk4\n'
            'controller = cfl\ndt = 0.01\ncfl = 0.4\n', 0.03),
}

for name, (case, n, kw, sect, tend) in INTG.items():
    kw2 = {k: v for k, v in kw.items() if k != 'warp'}
    txt = (cases.tgv_cfg(**kw2) if case == 'tgv' else cases.vortex_cfg(**kw2))
    txt += ('\n[backend-b200]\ngraphs = false\n[solver-time-integrator]\n'
            f'formulation = explicit\ntstart = 0\ntend = {tend}\n{sect}')
    _, box = cases.make(case, n, **kw)
    cls = NavierStokesSystem if case == 'tgv' else EulerSystem

    res = []
    for which in ('oracle', 'b200'):
        cfg = Inifile(txt)
        be = (B200Backend(cfg) if which == 'b200' else
              make_backend(rbase, name='oracle-ref')(cfg))
        intg = get_integrator(be, cls, rh.ref_mesh(box.local_mesh()), None,
                              cfg)
        intg.advance_to(tend)
        res.append((intg.soln[0].copy(), intg.nacptsteps, intg.nrjctsteps,
                    intg.dt))

    (so, ao, ro, dto), (sb, ab, rb, dtb) = res
    print('RESULT', name, np.abs(sb - so).max()/np.abs(so).max(), ao, ab, ro,
          rb, abs(dtb/dto - 1))
'''


@pytest.mark.skipif(not _ref_available(),
                    reason='needs /root/reference')
def test_reference_integrators_execute_on_b200_backend(built):
    """The reference's composed integrator classes (RK45 + PI controller,
    RK4 + CFL controller) drive B200Backend through their own call sites:
    ``kernel('rkvdh2', ...)``, ``kernel('reduction', rop, exprs, vvars,
    svars=, pvars=)`` with positional ``bind`` and ``retval``,
    ``kernel('wavespeed', ...)``, ``axnpby``.  Same decisions and solution
    as on the oracle backend."""
    res = subprocess.run([sys.executable, '-c',
                          _intg_script % {'root': ROOT}],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]

    rows = {l.split()[1]: l.split()[2:] for l in res.stdout.splitlines()
            if l.startswith('RESULT')}
    assert set(rows) == {'pi', 'cfl'}

    for name, (err, ao, ab, ro, rb, ddt) in rows.items():
        assert float(err) < 1e-12 and float(ddt) < 1e-9
        assert (ao, ro) == (ab, rb) and int(ao) >= 2

    assert int(rows['pi'][3]) >= 1                # a rejected step happened


_plugin_script = r'''
import os, sys, tempfile
os.environ['PYFR_B200_BASE'] = 'pyfr.backends.base'
ROOT = %(root)r
sys.path[:0] = [ROOT, ROOT + '/tests', ROOT + '/tests/cudaemu']
import numpy as np
from oracle import refharness as rh
from oracle.npbackend import LocalComm, make_backend
rh.install_stubs()
rh.set_rank(LocalComm(0, 1).peer(0))
import emu
import pyfr_b200.backend as bk, pyfr_b200.compiler as comp
bk.load_runtime = lambda device=0, dry=False: emu.EmuRuntime()
comp.KernelCompiler.cubin = lambda self, src, name: src.encode()
import pyfr.backends.base as rbase
from pyfr.inifile import Inifile
from pyfr.integrators import get_integrator
from pyfr.solvers.navstokes import NavierStokesSystem
from pyfr_b200 import cases
from pyfr_b200.backend import B200Backend
from pyfr_b200.host.integrator import TGV_EXPRS

_, box = cases.make('tgv', (3, 2, 2), order=2, warp=0.1)
tmp = tempfile.mkdtemp()
rows = {}
for which in ('oracle', 'b200'):
    csv = os.path.join(tmp, which + '.csv')
    txt = cases.tgv_cfg(order=2) + f"""
[backend-b200]
graphs = false
[solver-time-integrator]
formulation = explicit
scheme = rk4
controller = none
tstart = 0
tend = 0.006
dt = 0.002
[soln-plugin-integrate]
nsteps = 1
file = {csv}
header = true
int-ke = {TGV_EXPRS[0]}
int-ens = {TGV_EXPRS[1]}
[soln-plugin-integrate-linf]
nsteps = 1
file = {csv[:-4]}_linf.csv
header = true
norm = inf
int-q = rho*x*y - p*cos(z)
int-g = grad_u_y + 0.1*t
"""
    cfg = Inifile(txt)
    be = (B200Backend(cfg) if which == 'b200' else
          make_backend(rbase, name='oracle-ref')(cfg))
    intg = get_integrator(be, NavierStokesSystem,
                          rh.ref_mesh(box.local_mesh()), None, cfg)
    intg.advance_to(0.006)
    rows[which] = np.hstack([
        np.loadtxt(csv, delimiter=',', skiprows=1),
        np.loadtxt(csv[:-4] + '_linf.csv', delimiter=',', skiprows=1)[:, 1:]
    ])

print('RESULT', rows['oracle'].shape[0],
      np.abs(rows['b200']/rows['oracle'] - 1)[:, 1:].max(),
      rows['oracle'][0, 1]/(2*np.pi)**3)
'''


@pytest.mark.skipif(not _ref_available(),
                    reason='needs /root/reference')
def test_reference_integrate_plugin_on_b200_backend(built):
    """The reference's ``[soln-plugin-integrate]`` (``BackendFieldReducer``,
    ``compute_grads`` graph, ``fieldeval`` kernel) inside the reference's
    RK4 integrator, on B200Backend: the CSV of Taylor-Green kinetic energy
    and enstrophy equals the oracle backend's to round-off."""
    res = subprocess.run([sys.executable, '-c',
                          _plugin_script % {'root': ROOT}],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]

    _, nrows, err, ke0 = [l for l in res.stdout.splitlines()
                          if l.startswith('RESULT')][0].split()
    # columns: t, two volume integrals, two L-inf norms (max reduction,
    # coordinate- and time-dependent expressions)
    assert int(nrows) == 4 and float(err) < 1e-13
    assert abs(float(ke0) - 0.125) < 5e-3


_mpi_script = r'''
import ctypes as ct, os, sys
os.environ['PYFR_B200_BASE'] = 'pyfr.backends.base'
ROOT = %(root)r
sys.path[:0] = [ROOT, ROOT + '/tests', ROOT + '/tests/cudaemu']
from types import SimpleNamespace
import numpy as np
from oracle import refharness as rh
from oracle.npbackend import LocalComm
rh.install_stubs()
import emu
import pyfr_b200.backend as bk, pyfr_b200.compiler as comp
bk.load_runtime = lambda device=0, dry=False: emu.EmuRuntime()
comp.KernelCompiler.cubin = lambda self, src, name: src.encode()
from pyfr.inifile import Inifile
from pyfr.solvers.euler import EulerSystem
from pyfr.solvers.navstokes import NavierStokesSystem
from pyfr_b200 import cases
from pyfr_b200.backend import B200Backend
from test_emulated_kernels import EmuWorld

for case, n, parts, kw in [('tgv', (4, 2, 2), (2, 1, 1), dict(order=2, warp=0.1)),
                           ('tgv', (4, 2, 2), (2, 1, 1),
                            dict(order=2, beta=0.0, rsolver='hllc')),
                           ('vortex', (6, 4), (3, 1), dict(order=3))]:
    kw2 = {k: v for k, v in kw.items() if k != 'warp'}
    txt = (cases.tgv_cfg(**kw2) if case == 'tgv' else cases.vortex_cfg(**kw2))
    _, box = cases.make(case, n, **kw)
    nparts = int(np.prod(parts))
    vparts = box.brick_partition(parts)
    cls = NavierStokesSystem if case == 'tgv' else EulerSystem

    # The reference's host code per rank: on the oracle backend ...
    lworld = LocalComm(0, nparts)
    ref = [rh.ref_system(txt, box.local_mesh(vparts, r), 2, lworld.peer(r))[0]
           for r in range(nparts)]

    # ... and on the B200 backend with an in-process communicator
    eworld = EmuWorld(nparts)
    b200 = []
    for r in range(nparts):
        rh.set_rank(lworld.peer(r))
        cfg = Inifile(txt + '\n[backend-b200]\ngraphs = true\n')
        comm = eworld.peer(r)
        be = B200Backend(cfg, comm=comm)
        comm.rt = be.rt
        regs = [SimpleNamespace(rhs=True, dynamic=False, n=2, extent=None)]
        s = cls(be, rh.ref_mesh(box.local_mesh(vparts, r)), None, regs, cfg,
                None)
        s.commit()
        b200.append(s)

    for systems, world in ((ref, lworld), (b200, eworld)):
        graphs = [s._rhs_graphs(0, 1) for s in systems]
        for s in systems:
            s._prepare_kernels(0.0, 0, 1)
        for stage in zip(*graphs):
            for g in stage:
                g.run()
            world.deliver()

    err = max(np.abs(a.ele_scal_upts(1)[0] - b.ele_scal_upts(1)[0]).max()
              / np.abs(a.ele_scal_upts(1)[0]).max()
              for a, b in zip(ref, b200))
    nx = sum(1 for g in b200[0]._rhs_graphs(0, 1)
             for w, o in g.plan if w == 'xchg')
    print('RESULT', case, nparts, err, nx)
'''


@pytest.mark.skipif(not _ref_available(),
                    reason='needs /root/reference')
def test_reference_mpi_interfaces_on_b200_backend(built):
    """Partitioned runs under the reference's own host code: its
    ``MPIInters`` classes create the exchange views, pack kernels and
    ``sendreq``/``recvreq`` requests (``register_mpi_exchange``,
    pyfr/solvers/base/system.py:185-202) and hand them to ``Graph.
    add_mpi_req``; the backend turns them into grouped exchanges inside
    captured graphs.  Two and three ranks, same RHS as the reference host on
    the oracle backend."""
    res = subprocess.run([sys.executable, '-c',
                          _mpi_script % {'root': ROOT}],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2500:]

    rows = [l.split()[1:] for l in res.stdout.splitlines()
            if l.startswith('RESULT')]
    assert [(r[0], int(r[1])) for r in rows] == [('tgv', 2), ('tgv', 2),
                                                 ('vortex', 3)]
    assert all(float(r[2]) < 1e-12 for r in rows)
    # Navier-Stokes exchanges twice per RHS, Euler once
    assert [int(r[3]) for r in rows] == [2, 2, 1]


_bc_script = r'''
import os, sys
os.environ['PYFR_B200_BASE'] = 'pyfr.backends.base'
ROOT = %(root)r
sys.path[:0] = [ROOT, ROOT + '/tests', ROOT + '/tests/cudaemu']
from types import SimpleNamespace
import numpy as np
from oracle import refharness as rh
from oracle.npbackend import LocalComm
rh.install_stubs()
import emu
import pyfr_b200.backend as bk, pyfr_b200.compiler as comp
bk.load_runtime = lambda device=0, dry=False: emu.EmuRuntime()
comp.KernelCompiler.cubin = lambda self, src, name: src.encode()
from pyfr.inifile import Inifile
from pyfr.solvers.euler import EulerSystem
from pyfr.solvers.navstokes import NavierStokesSystem
from pyfr_b200 import cases
from pyfr_b200.backend import B200Backend

CASES = [
    ('navier-stokes', (3, 3, 2), {'ylo': 'no-slp-adia-wall',
                                  'yhi': 'char-riem-inv',
                                  'xlo': 'sub-in-ftpttang',
                                  'xhi': 'sub-out-fp'},
     dict(order=2, warp=0.1)),
    ('navier-stokes', (2, 2, 3), {'xlo': 'sub-in-frv', 'xhi': 'sup-out-fn',
                                  'zlo': 'slp-adia-wall',
                                  'zhi': 'no-slp-isot-wall',
                                  'ylo': 'sup-in-fa', 'yhi': 'sub-out-fp'},
     dict(order=2, rsolver='hllc', beta=0.0)),
    ('euler', (5, 4), {'xlo': 'char-riem-inv', 'xhi': 'sup-out-fn',
                       'ylo': 'slp-adia-wall', 'yhi': 'sup-in-fa'},
     dict(order=3)),
]

for system, n, bcs, kw in CASES:
    _, box, txt = cases.box_case(system, n, bcs, **kw)
    mesh = box.local_mesh()
    world = LocalComm(0, 1)
    rh.set_rank(world.peer(0))

    cfg = Inifile(txt + '\n[backend-b200]\ngraphs = true\n')
    be = B200Backend(cfg)
    regs = [SimpleNamespace(rhs=True, dynamic=False, n=2, extent=None)]
    cls = NavierStokesSystem if system == 'navier-stokes' else EulerSystem
    s = cls(be, rh.ref_mesh(mesh), None, regs, cfg, None)
    s.commit()

    rs, rbe = rh.ref_system(txt, mesh, 2, world.peer(0))
    errs = []
    for t in (0.0, 0.7):
        s.rhs(t, 0, 1)
        rs.rhs(t, 0, 1)
        out, ref = s.ele_scal_upts(1)[0], rs.ele_scal_upts(1)[0]
        errs.append(np.abs(out - ref).max()/np.abs(ref).max())
    print('RESULT', system, len(bcs), max(errs))
'''


@pytest.mark.skipif(not _ref_available(),
                    reason='needs /root/reference')
def test_reference_boundary_interfaces_on_b200_backend(built):
    """The reference's boundary-interface classes (``pyfr/solvers/*/
    inters.py``: template arguments, boundary-section constants and
    expressions, ``ploc`` views) drive ``bcconu`` / ``bccflux`` for all
    nine boundary types on the path."""
    res = subprocess.run([sys.executable, '-c', _bc_script % {'root': ROOT}],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2500:]

    rows = [l.split()[1:] for l in res.stdout.splitlines()
            if l.startswith('RESULT')]
    assert len(rows) == 3 and all(float(r[2]) < 1e-12 for r in rows)


_device_script = r"""
import os, sys
os.environ['PYFR_B200_BASE'] = 'pyfr.backends.base'
ROOT = %(root)r
sys.path[:0] = [ROOT, ROOT + '/tests']
from types import SimpleNamespace
import numpy as np
from oracle import refharness as rh
from oracle.npbackend import LocalComm
rh.install_stubs()
from pyfr.inifile import Inifile
from pyfr.backends.base import BaseBackend
from pyfr.solvers.navstokes import NavierStokesSystem
from pyfr.solvers.euler import EulerSystem
from pyfr.util import subclass_where
from pyfr_b200 import cases
from pyfr_b200.backend import B200Backend

# the reference's own backend discovery finds the backend by name
# (pyfr/backends/__init__.py:10-11)
assert subclass_where(BaseBackend, name='b200') is B200Backend
assert issubclass(B200Backend, BaseBackend)

for case, n, kw in [('tgv', (4, 3, 3), dict(order=4)),
                    ('tgv', (4, 3, 3), dict(order=3, warp=0.1,
                                            rsolver='hllc', beta=0.0)),
                    ('vortex', 8, dict(order=3))]:
    kw2 = {k: v for k, v in kw.items() if k != 'warp'}
    txt = (cases.tgv_cfg(**kw2) if case == 'tgv' else cases.vortex_cfg(**kw2))
    _, box = cases.make(case, n, **kw)
    mesh = box.local_mesh()
    world = LocalComm(0, 1)
    rh.set_rank(world.peer(0))

    be = B200Backend(Inifile(txt))
    assert not be.rt.dry and be.use_graphs
    regs = [SimpleNamespace(rhs=True, dynamic=False, n=2, extent=None)]
    cls = NavierStokesSystem if case == 'tgv' else EulerSystem
    s = cls(be, rh.ref_mesh(mesh), None, regs, Inifile(txt), None)
    s.commit()
    for _ in range(2):
        s.rhs(0.0, 0, 1)
    be.wait()
    out = s.ele_scal_upts(1)[0]
    kinds = [getattr(k, 'kind', None) or k.fn.name
             for g in s._rhs_graphs(0, 1) for w, k in g.plan if w == 'kernel']

    # reference host code on the NumPy oracle backend, same mesh
    rs, rbe = rh.ref_system(txt, mesh, 2, world.peer(0))
    rs.rhs(0.0, 0, 1)
    ref = rs.ele_scal_upts(1)[0]
    print('RESULT', case, np.abs(out - ref).max()/np.abs(ref).max(),
          ','.join(kinds))
"""


@pytest.mark.gpu
@pytest.mark.skipif(not _ref_available(),
                    reason='no PyFR tree (PYFR_B200_REFROOT, /root/reference '
                           'or baseline/_ref)')
def test_reference_host_drives_the_device(built):
    """The drop-in on hardware: the reference's unmodified
    ``NavierStokesSystem`` / ``EulerSystem`` (from the PyFR tree that is
    present) drive ``B200Backend`` -- derived from the reference's own
    ``pyfr.backends.base`` classes and found by its backend discovery -- on
    the device, CUDA graphs on; the RHS equals the reference host on the
    NumPy oracle backend."""
    res = subprocess.run([sys.executable, '-c',
                          _device_script % {'root': ROOT}],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]

    rows = [l.split() for l in res.stdout.splitlines()
            if l.startswith('RESULT')]
    assert [r[1] for r in rows] == ['tgv', 'tgv', 'vortex']
    # (the p = 4 case is held to the oracle's fp64 floor at that order)
    assert float(rows[0][2]) < 3e-11 and float(rows[1][2]) < 5e-12
    assert float(rows[2][2]) < 1e-12
    assert rows[0][3].split(',') == ['mul', 'gradflux', 'intcflux',
                                     'mul+negdivconf']
    assert 'fluxdiv' in rows[2][3]

    from util import PARITY_LOG
    for r in rows:
        PARITY_LOG.append(dict(test=f'reference host on the device: {r[1]}',
                               err=float(r[2]), floor=0.0, ratio=None,
                               ratio_oracle=None, kernels=r[3]))
