"""TEST ORACLE -- drive the reference's *own* host code in this container.

Only usable where a reference tree exists (``/root/reference`` in the build
container, or the offline install ``baseline/_ref``); never imported by the
product, by ``smoke()`` or by ``bench.py``.  The one ``-m gpu`` test that
uses it (tests/test_reference_dropin.py) lets the reference's *host code*
drive the B200 backend on the device and skips where no tree is present.

The reference's solver stack imports a handful of third-party packages that
are absent here (mako, mpi4py, h5py, rtree, pytools, gimmik).  None of them
is *executed* on the path we need -- element geometry, connectivity, view
construction, graph assembly -- so name-only stubs are enough for
``pyfr.solvers`` to import and for the real ``NavierStokesSystem`` /
``EulerSystem`` to run on top of the NumPy oracle backend.  This is how the
host mirror in ``pyfr_b200/host`` is pinned: same mesh + same backend, the
reference's host code and ours must produce identical view indices and RHS.
"""

import ctypes.util
import os
import sys
import types
from types import SimpleNamespace

import numpy as np

def _find_reference():
    """The reference tree: ``$PYFR_B200_REFROOT``, the read-only checkout of
    the build container, or the offline install ``baseline/_ref`` (``pip
    install --no-deps --target baseline/_ref``; git-ignored, but it travels
    with the repository snapshot to a GPU box)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cands = [os.environ.get('PYFR_B200_REFROOT'), '/root/reference',
             os.path.join(here, 'baseline', '_ref')]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, 'pyfr')):
            return c
    return cands[0] or cands[1]


REFROOT = _find_reference()


def available():
    return os.path.isdir(os.path.join(REFROOT, 'pyfr'))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class StubComm:
    """Stands in for MPI.COMM_WORLD; point-to-point goes through a
    oracle.npbackend.LocalComm so several ranks can share one process."""

    def __init__(self, local):
        self.local = local
        self.rank, self.size = local.rank, local.size

    def send_init(self, xm, pid, tag):
        return self.local.send_init(xm, pid, tag)

    def recv_init(self, xm, pid, tag):
        return self.local.recv_init(xm, pid, tag)

    # Scalar collectives of the integrators (pyfr/mpiutil.py:99-103); the
    # harness only ever reduces over a single rank
    def Allreduce(self, sendbuf, recvbuf, op=None):
        if self.size != 1:
            raise NotImplementedError('multi-rank reference integrator')

    def Reduce(self, sendbuf, recvbuf, op=None, root=0):
        if self.size != 1:
            raise NotImplementedError('multi-rank reference plugins')

    def exscan(self, v):
        return None                     # rank 0's result is undefined in MPI

    # Per-boundary communicators (pyfr/solvers/base/system.py:275-278);
    # only their existence matters on this path
    handle = 0

    def Split(self, color, key=0):
        return self

    @staticmethod
    def fromhandle(handle):
        return SimpleNamespace(free=lambda: None)


def install_stubs():
    if 'pyfr' in sys.modules:
        return

    if REFROOT not in sys.path:
        sys.path.insert(0, REFROOT)

    class _Any:
        def __init__(self, *a, **k): pass
        def __getattr__(self, n): return _Any()
        def __call__(self, *a, **k): return _Any()

    _mod('mako')
    _mod('mako.runtime', supports_caller=lambda f: f, capture=None)
    _mod('mako.lookup', TemplateLookup=_Any)
    _mod('mako.template', Template=_Any)
    _mod('h5py', File=_Any, Dataset=_Any, Group=_Any)
    _mod('rtree')
    _mod('rtree.index', Index=_Any, Property=_Any)
    _mod('pytools')
    _mod('pytools.prefork', enable_prefork=lambda: None, call_capture_output=None)
    _mod('gimmik')

    class Prequest:
        @staticmethod
        def Startall(reqs): pass

        @staticmethod
        def Waitall(reqs): pass

    mpi4py = _mod('mpi4py')
    mpi4py.rc = _mod('mpi4py.rc')
    mpi4py.MPI = _mod(
        'mpi4py.MPI', __file__=ctypes.util.find_library('c') or 'libc.so.6',
        Prequest=Prequest, COMM_WORLD=None, MIN=0, MAX=1, SUM=2, LOR=3,
        IN_PLACE=4, UNDEFINED=-1
    )


def set_rank(local):
    import mpi4py.MPI as MPI
    MPI.COMM_WORLD = StubComm(local)


def ref_mesh(m):
    """Our in-memory mesh -> the reference's Mesh dataclass."""
    from pyfr.readers.native import Connectivity, Mesh

    con = lambda c: Connectivity(np.asarray(c.cidxs), np.asarray(c.eidxs),
                                 c.cidxmap)

    return Mesh(
        fname='synthetic', raw=None, ndims=m.ndims, codec=list(m.codec),
        uuid=m.uuid, etypes=list(m.etypes), eidxs=dict(m.eidxs),
        spts=dict(m.spts), spts_curved=dict(m.spts_curved),
        con=tuple(con(c) for c in m.con),
        con_p={p: con(c) for p, c in m.con_p.items()},
        bcon={b: con(c) for b, c in m.bcon.items()}, cidxmap=m.cidxmap
    )


def ref_system(cfgtext, mesh, nregs, local):
    """Build the reference's system class on the oracle backend."""
    install_stubs()
    set_rank(local)

    import pyfr.backends.base as rbase
    from pyfr.inifile import Inifile
    from pyfr.solvers.euler import EulerSystem
    from pyfr.solvers.navstokes import NavierStokesSystem

    from oracle.npbackend import make_backend

    cfg = Inifile(cfgtext)
    be = make_backend(rbase, name='oracle-ref')(cfg)

    regs = [SimpleNamespace(rhs=True, dynamic=False, n=nregs, extent=None)]
    cls = {'euler': EulerSystem,
           'navier-stokes': NavierStokesSystem}[cfg.get('solver', 'system')]

    system = cls(be, ref_mesh(mesh), None, regs, cfg, None)
    system.commit()

    return system, be
