"""BASELINE.json configs[3]: mixed element types (quad + tri; hex + pri +
pyr + tet) through the B200 backend.

The host mirror in ``pyfr_b200/host`` only carries the tensor-product
shapes, so here the *reference's own* solver classes (``pyfr.solvers``,
``pyfr.shapes``: Williams-Shunn / Shunn-Ham point sets, dense operators,
mixed-face interface views spanning several element types) drive
``B200Backend`` -- derived from the reference's base classes -- with the
generated CUDA kernels executed on the CPU model of tests/cudaemu, and the
NumPy oracle backend beside it.  Needs /root/reference; the conforming
meshes come from tests/mixedmesh.py.  Every generated kernel source is also
compiled for sm_100a with nvcc."""

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

pytestmark = pytest.mark.skipif(
    not _ref_available(), reason='needs /root/reference'
)

_script = r'''
import os, re, sys
os.environ['PYFR_B200_BASE'] = 'pyfr.backends.base'
ROOT = %(root)r
sys.path[:0] = [ROOT, ROOT + '/tests', ROOT + '/tests/cudaemu']
from concurrent.futures import ThreadPoolExecutor
from types import SimpleNamespace
import numpy as np
from oracle import refharness as rh
from oracle.npbackend import LocalComm, make_backend
rh.install_stubs()
rh.set_rank(LocalComm(0, 1).peer(0))
import emu
import pyfr_b200.backend as bk, pyfr_b200.compiler as comp
sources = {}
def cubin(self, src, name):
    sources[src] = name
    return src.encode()
bk.load_runtime = lambda device=0, dry=False: emu.EmuRuntime()
comp.KernelCompiler.cubin = cubin
import pyfr.backends.base as rbase
from pyfr.inifile import Inifile
from pyfr.solvers.euler import EulerSystem
from pyfr.solvers.navstokes import NavierStokesSystem
from pyfr_b200 import cases
from pyfr_b200.backend import B200Backend
import mixedmesh as mm

POINTS = """
[solver-interfaces-line]
flux-pts = gauss-legendre
[solver-interfaces-quad]
flux-pts = gauss-legendre
[solver-interfaces-tri]
flux-pts = williams-shunn
[solver-elements-quad]
soln-pts = gauss-legendre
[solver-elements-tri]
soln-pts = williams-shunn
[solver-elements-hex]
soln-pts = gauss-legendre
[solver-elements-tet]
soln-pts = shunn-ham
[solver-elements-pri]
soln-pts = williams-shunn~gauss-legendre
[solver-elements-pyr]
soln-pts = gauss-legendre
[backend-b200]
graphs = false
"""

ICS2 = ('[soln-ics]\nrho = 1 + 0.1*sin(kx*x)*cos(ky*y)\n'
        'u = 0.3 + 0.1*cos(kx*x + ky*y)\nv = 0.15 + 0.1*sin(ky*y)\n'
        'p = 4.5*(1 + 0.02*cos(kx*x))\n')
ICS3 = ('[soln-ics]\nrho = 1 + 0.1*sin(kx*x)*cos(ky*y)\n'
        'u = 0.3 + 0.1*cos(kx*x + ky*y)\nv = 0.15 + 0.1*sin(ky*y)*cos(kz*z)\n'
        'w = 0.1 + 0.05*sin(kz*z + kx*x)\n'
        'p = 71*(1 + 0.02*cos(kx*x)*sin(kz*z))\n')

CASES = {
    'quad+tri': (mm.columns(4, 3, None, ['quad', 'tri']),
                 cases.vortex_cfg(order=3, rsolver='hllc'), EulerSystem),
    'hex+pri': (mm.columns(3, 2, 2, ['hex', 'pri']),
                cases.tgv_cfg(order=2, beta=0.0), NavierStokesSystem),
    'hex+pri+pyr+tet': (mm.columns(4, 2, 2, ['hex', 'pri', 'pyr', 'pyt']),
                        cases.tgv_cfg(order=3), NavierStokesSystem),
}

for name in %(names)r:
    kinds, txt, cls = CASES[name]
    ks = ''.join(f'k{"xyz"[a]} = {2*np.pi/kinds.shape[a]!r}\n'
                 for a in range(kinds.ndim))
    head = txt.partition('[soln-ics]')[0]
    head = head.replace('[constants]\n', '[constants]\n' + ks)
    extra = POINTS
    for sect in re.findall(r'^\[([^\]]+)\]', head, flags=re.M):
        extra = re.sub(r'\[' + re.escape(sect) + r'\]\n[^\[]*', '', extra)
    txt = head + (ICS2 if kinds.ndim == 2 else ICS3) + extra

    mesh = mm.build(kinds, h=1.0, warp=0.05)
    outs = []
    for which in ('oracle', 'b200'):
        cfg = Inifile(txt)
        be = (B200Backend(cfg) if which == 'b200' else
              make_backend(rbase, name='oracle-ref')(cfg))
        regs = [SimpleNamespace(rhs=True, dynamic=False, n=2, extent=None)]
        s = cls(be, mesh, None, regs, cfg, None)
        s.commit()
        s.rhs(0.0, 0, 1)
        outs.append(s.ele_scal_upts(1))
        if which == 'b200':
            kk = [getattr(k, 'kind', None) for g in s._rhs_graphs(0, 1)
                  for w, k in g.plan if w == 'kernel']
            print('KINDS', name, ' '.join(map(str, kk)))

    for et, a, b in zip(mesh.etypes, *outs):
        print('RESULT', name, et, a.shape[0], a.shape[2],
              np.abs(a - b).max()/np.abs(a).max())

if %(compile)r:
    def one(item):
        src, name = item
        image, log = comp.compile_nvcc(src, name)
        return name, len(image)
    with ThreadPoolExecutor(8) as ex:
        for name, n in ex.map(one, sources.items()):
            print('COMPILED', name, n)
'''


def _run(names, compile=False):
    res = subprocess.run(
        [sys.executable, '-c',
         _script % {'root': ROOT, 'names': names, 'compile': compile}],
        capture_output=True, text=True, timeout=1500
    )
    assert res.returncode == 0, res.stderr[-3000:]

    rows = [l.split()[1:] for l in res.stdout.splitlines()
            if l.startswith('RESULT')]
    kinds = {l.split()[1]: l.split()[2:] for l in res.stdout.splitlines()
             if l.startswith('KINDS')}
    comp = [l.split()[1:] for l in res.stdout.splitlines()
            if l.startswith('COMPILED')]
    return rows, kinds, comp


def test_euler_on_quads_and_triangles(built):
    rows, kinds, _ = _run(['quad+tri'])

    assert [(r[1], int(r[2])) for r in rows] == [('quad', 16), ('tri', 10)]
    assert all(float(r[4]) < 1e-12 for r in rows)
    # both element types take the fused flux-divergence kernel
    assert kinds['quad+tri'].count('fluxdiv') == 2


def test_navier_stokes_on_hexes_and_prisms(built):
    rows, kinds, _ = _run(['hex+pri'])

    assert [(r[1], int(r[2])) for r in rows] == [('hex', 27), ('pri', 18)]
    assert all(float(r[4]) < 1e-12 for r in rows)
    assert kinds['hex+pri'].count('gradflux') == 2


def test_navier_stokes_on_all_3d_element_types(built):
    """configs[3]: order 3, hexes, prisms, pyramids and tetrahedra in one
    mesh.  The dense tet / pyramid operators have no line structure, so
    those types fall back to the individual kernels (an in-place divergence
    over overlapping row groups would be wrong -- regression test)."""
    rows, kinds, comp = _run(['hex+pri+pyr+tet'], compile=True)

    assert [(r[1], int(r[2])) for r in rows] == [('hex', 64), ('pri', 40),
                                                 ('pyr', 30), ('tet', 20)]
    assert all(float(r[4]) < 1e-12 for r in rows)

    kk = kinds['hex+pri+pyr+tet']
    assert kk.count('gradflux') == 2 and kk.count('tflux') == 2
    assert kk.count('mul+negdivconf') == 4

    # every generated source also builds for sm_100a
    assert len(comp) >= 20 and all(int(n) > 0 for _, n in comp)


@pytest.mark.parametrize('args,parts', [
    ((4, 4, None, ['quad', 'tri']), (2, 2)),
    ((4, 2, 2, ['hex', 'pri', 'pyr', 'pyt']), (2, 1, 1)),
    ((4, 4, 2, ['hex', 'pri', 'pyr', 'pyt']), (2, 2, 1)),
], ids=str)
def test_partitioned_mixed_mesh_matches_reference_reader(args, parts):
    """Brick-partitioned mixed meshes: element order, interior and
    inter-partition connectivity of every rank equal what the reference's
    reader derives from the same per-face neighbour records (bit-exact)."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    sys.path.insert(0, ROOT)
    import mixedmesh as mm
    from pyfr_b200.host.mesh import MixedBoxMesh

    box = MixedBoxMesh(MixedBoxMesh.columns(*args), h=1.0, warp=0.05)
    vparts = box.brick_partition(parts)
    ref = mm.ref_partitioned_con(box, vparts)

    assert len(ref) == int(np.prod(parts))
    for r, rm in enumerate(ref):
        m = box.local_mesh(vparts, r)

        assert set(m.eidxs) == set(rm.eidxs)
        for a, b in zip(m.con, rm.con):
            assert np.array_equal(a.cidxs, b.cidxs)
            assert np.array_equal(a.eidxs, b.eidxs)

        assert set(m.con_p) == set(rm.con_p) and m.con_p
        for p in m.con_p:
            assert np.array_equal(m.con_p[p].cidxs, rm.con_p[p].cidxs)
            assert np.array_equal(m.con_p[p].eidxs, rm.con_p[p].eidxs)
