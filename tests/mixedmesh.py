"""Synthetic conforming mixed-element boxes (BASELINE.json configs[3]) in
the reference's in-memory ``Mesh`` form.  TEST INFRASTRUCTURE: needs
``/root/reference`` (shape tables, ``NativeReader._construct_con``), so it
is only used by tests that are skipped where the reference is absent.

Layout: a periodic box of ``n`` cells per
direction; every column of cells (all z) has one kind:

  2-D  ``quad`` | ``tri``  (cell split on the anti-diagonal)
  3-D  ``hex``  | ``pri``  (that triangle pair extruded through the cell)
       ``pyr``  (six pyramids, apex at the cell centre)
       ``pyt``  (as ``pyr`` with the top and bottom pyramids split into two
                 tetrahedra each, on the same diagonal)

Neighbouring columns always meet in whole quadrilateral faces and cells
stacked in z share their kind, so the mesh is conforming.  Faces are paired
by their centroids (modulo the period); the interior connectivity is then
derived by the reference's own reader code.
"""

from types import SimpleNamespace

import numpy as np

from oracle import refharness as rh

# Vertex lists in the reference's std-element order (pyfr/shapes.py
# std_ele(1)), as corner keys (dx, dy[, dz]) of the unit cell; 'P' = centre
_Q = {'quad': [[(0, 0), (1, 0), (0, 1), (1, 1)]],
      'tri': [[(0, 0), (1, 0), (0, 1)], [(1, 1), (0, 1), (1, 0)]]}

_PYR_BASES = [
    [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0)],       # z lo
    [(0, 0, 1), (0, 1, 1), (1, 0, 1), (1, 1, 1)],       # z hi
    [(0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1)],       # x lo
    [(1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)],       # x hi
    [(0, 0, 0), (0, 0, 1), (1, 0, 0), (1, 0, 1)],       # y lo
    [(0, 1, 0), (1, 1, 0), (0, 1, 1), (1, 1, 1)],       # y hi
]


def _cell_elements(kind):
    """{etype: [vertex key lists]} for one cell of the given kind."""
    if kind in ('quad', 'tri'):
        return {kind: _Q[kind]}
    if kind == 'hex':
        return {'hex': [[(i, j, k) for k in (0, 1) for j in (0, 1)
                         for i in (0, 1)]]}
    if kind == 'pri':
        return {'pri': [[v + (k,) for k in (0, 1) for v in t]
                        for t in _Q['tri']]}
    if kind == 'pyr':
        return {'pyr': [b + ['P'] for b in _PYR_BASES]}
    if kind == 'pyt':
        tets = []
        for v0, v1, v2, v3 in _PYR_BASES[:2]:
            tets += [[v0, v1, v3, 'P'], [v0, v3, v2, 'P']]
        return {'pyr': [b + ['P'] for b in _PYR_BASES[2:]], 'tet': tets}
    raise ValueError(kind)


def build(kinds, h=1.0, warp=0.0):
    """``kinds``: array of cell kinds, shape ``(nx, ny)`` in 2-D or
    ``(nx, ny, nz)`` in 3-D (z-invariant per the module docstring).
    Returns the reference ``Mesh`` of the fully periodic box."""
    rh.install_stubs()
    import pyfr.readers.native as rnative
    from pyfr.polys import get_polybasis
    from pyfr.readers.native import Mesh
    from pyfr.shapes import BaseShape
    from pyfr.util import subclass_where

    kinds = np.asarray(kinds)
    nd, n = kinds.ndim, kinds.shape
    L = np.array(n, dtype=float)*h

    # -- elements ------------------------------------------------------------
    verts = {}
    for idx in np.ndindex(*n):
        org = np.array(idx, dtype=float)
        for et, lists in _cell_elements(kinds[idx]).items():
            for vl in lists:
                pts = [(org + 0.5 if v == 'P' else org + np.array(v))*h
                       for v in vl]
                verts.setdefault(et, []).append(pts)

    etypes = sorted(verts)
    spts = {et: np.array(verts[et]).swapaxes(0, 1) for et in etypes}

    # -- face centroids --------------------------------------------------------
    fcent = {'line': (0.0,), 'quad': (0.0, 0.0), 'tri': (-1/3, -1/3)}
    codec = [f'eles/{et}' for et in etypes]
    table = {}                                  # key -> [(etype, ele, face)]

    for et in etypes:
        scls = subclass_where(BaseShape, name=et)
        sord = scls.order_from_npts(len(spts[et]))
        sbasis = get_polybasis(et, sord, scls.std_ele(sord))
        for fidx, (ftype, proj, _) in enumerate(scls.faces):
            codec.append(f'eles/{et}/face/{fidx}')
            op = sbasis.nodal_basis_at([proj(*fcent[ftype])])
            cen = np.einsum('ij,jek->ek', op, spts[et])
            keys = np.rint(np.mod(cen, L)*12/h).astype(int) % \
                (12*np.array(n))
            for e, k in enumerate(map(tuple, keys)):
                table.setdefault(k, []).append((et, e, fidx))

    if any(len(v) != 2 for v in table.values()):
        raise RuntimeError('non-conforming mixed mesh')

    nfaces = {et: len(subclass_where(BaseShape, name=et).faces)
              for et in etypes}
    faces = {et: np.zeros((spts[et].shape[1], nfaces[et]),
                          dtype=[('cidx', np.int16), ('off', np.int64)])
             for et in etypes}
    for (a, b) in table.values():
        for (et, e, f), (net, ne, nf) in ((a, b), (b, a)):
            faces[et][e, f] = (codec.index(f'eles/{net}/face/{nf}'), ne)

    # -- smooth warp (keeps periodicity and conformity) ---------------------
    if warp:
        for et in etypes:
            x = spts[et]
            ph = 2*np.pi*x/L
            spts[et] = x + warp*h*np.stack(
                [np.sin(ph[..., (a + 1) % nd] + 0.5 + 0.4*a)
                 for a in range(nd)], axis=-1
            )

    # -- interior connectivity by the reference's reader ---------------------
    rd = rnative.NativeReader.__new__(rnative.NativeReader)
    rd.mesh = SimpleNamespace(
        codec=codec, etypes=etypes, bcon={}, con_p={},
        eidxs={et: np.arange(spts[et].shape[1]) for et in etypes}
    )
    rd.eles = {et: {'faces': faces[et]} for et in etypes}
    rd.f = {f'eles/{et}': np.empty(spts[et].shape[1]) for et in etypes}
    rd.neighbours = []
    rd._construct_con()
    m = rd.mesh

    return Mesh(
        fname='synthetic-mixed', raw=None, ndims=nd, codec=codec,
        uuid='mixed', etypes=etypes, eidxs=dict(m.eidxs), spts=spts,
        spts_curved={et: np.zeros(spts[et].shape[1], dtype=bool)
                     for et in etypes},
        con=m.con, con_p={}, bcon={}, cidxmap=m.cidxmap
    )


def columns(nx, ny, nz, pattern):
    """Cell kinds for an ``nx x ny (x nz)`` box; ``pattern[(i + 2 j) %
    len(pattern)]`` picks the kind of column (i, j)."""
    k2 = np.array([[pattern[(i + 2*j) % len(pattern)] for j in range(ny)]
                   for i in range(nx)], dtype=object)
    return k2 if nz is None else np.repeat(k2[:, :, None], nz, axis=2)


def ref_partitioned_con(box, vparts):
    """Interior and inter-partition connectivity of every rank of a
    partitioned ``pyfr_b200.host.mesh.MixedBoxMesh``, derived by the
    reference's ``NativeReader._construct_con`` -- every rank in a thread
    with an in-process stand-in for the MPI neighbourhood collectives (as
    tests/golden/make_golden.py does for the hex boxes).  Returns one
    namespace per rank with ``eidxs``, ``con`` and ``con_p``."""
    import threading

    rh.install_stubs()
    import pyfr.readers.native as rnative

    order = box.partition_order(vparts)
    nparts = len(order)
    etof = {c: et for c, (et, f) in box.cidxmap.items()}

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

    orig = rnative.get_comm_rank_root
    rnative.get_comm_rank_root = lambda: (Comm(), tls.rank, 0)

    results, errors = {}, []

    def run(rank):
        try:
            tls.rank = rank
            gidx = order[rank]
            eles, nbrs = {}, set()

            for et, g in gidx.items():
                fc = box.faces[et][g]
                faces = np.empty(fc.shape[:2], dtype=[('cidx', np.int16),
                                                      ('off', np.int64)])
                faces['cidx'], faces['off'] = fc[..., 0], fc[..., 1]
                eles[et] = {'faces': faces}

                for c in np.unique(fc[..., 0]):
                    sel = fc[..., 0] == c
                    nbrs |= set(vparts[etof[c]][fc[..., 1][sel]].tolist())

            rd = rnative.NativeReader.__new__(rnative.NativeReader)
            rd.mesh = SimpleNamespace(codec=list(box.codec),
                                      etypes=list(box.etypes),
                                      eidxs=dict(gidx), bcon={}, con_p={})
            rd.eles = eles
            rd.f = {f'eles/{et}': np.empty(box._x0[et].shape[1])
                    for et in box.etypes}
            rd.neighbours = sorted(nbrs - {rank})
            rd._construct_con()
            results[rank] = rd.mesh
        except Exception as e:                          # pragma: no cover
            errors.append(e)
            barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(nparts)]
    try:
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    finally:
        rnative.get_comm_rank_root = orig

    if errors:
        raise errors[0]

    return [results[r] for r in range(nparts)]
