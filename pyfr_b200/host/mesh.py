"""Synthetic structured meshes in the in-memory form the solvers consume.

The reference reads a ``.pyfrm`` HDF5 file into a ``Mesh`` dataclass
(``pyfr/readers/native.py:15-46``) and derives the interior / boundary /
inter-partition connectivity in ``NativeReader._construct_con``
(``:445-534``).  Test-case meshes are not available offline, so this module
generates periodic (or walled) quad/hex boxes programmatically and applies
the same ordering rules, which makes every index downstream reproducible:

* elements inside a partition: by element type, partition-boundary
  elements first, then global number (``pyfr/partitioners/base.py:286-290``);
* faces are visited type by type, face by face, element by element; an
  interior face is listed once with the smaller ``(cidx, element)`` key on
  the left (``native.py:432-478``);
* inter-partition faces towards one neighbour are ordered by the
  ``(cidx, global element)`` key of the higher-ranked side (``:518-521``).
"""

from dataclasses import dataclass, field

import numpy as np

from pyfr_b200.host.shapes import shape_map


class Connectivity:
    """One side of a list of interfaces: codec index + element per face."""

    def __init__(self, cidxs, eidxs, cidxmap):
        self.cidxs = np.asarray(cidxs, dtype=np.int16)
        self.eidxs = np.asarray(eidxs, dtype=np.int64)
        self.cidxmap = cidxmap
        self._ucidxs = np.unique(self.cidxs).tolist()

    def __len__(self):
        return len(self.cidxs)

    def items(self):
        for c in self._ucidxs:
            etype, fidx = self.cidxmap[c]
            yield etype, fidx, self.eidxs[self.cidxs == c]

    def foreach(self):
        for c in self._ucidxs:
            sel = self.cidxs == c
            etype, fidx = self.cidxmap[c]
            yield etype, fidx, self.eidxs[sel], np.flatnonzero(sel)


@dataclass
class Mesh:
    ndims: int
    codec: list
    etypes: list
    cidxmap: dict
    uuid: str = 'synthetic'
    eidxs: dict = field(default_factory=dict)
    spts: dict = field(default_factory=dict)
    spts_curved: dict = field(default_factory=dict)
    con: tuple = ()
    con_p: dict = field(default_factory=dict)
    bcon: dict = field(default_factory=dict)


# Face pairing of the tensor-product reference elements: face f of an
# element meets face opp[f] of its neighbour in direction (axis, sign)
_face_dirs = {
    'quad': [(1, -1), (0, 1), (1, 1), (0, -1)],
    'hex': [(2, -1), (1, -1), (0, 1), (1, 1), (0, -1), (2, 1)],
}
_face_opp = {'quad': [2, 3, 0, 1], 'hex': [5, 3, 4, 1, 2, 0]}


class BoxMesh:
    """Global description of an ``n[0] x n[1] (x n[2])`` box of quads/hexes.

    Element ``g = i + n0*(j + n1*k)``; periodic axes wrap, other axes end in
    boundaries named ``'<axis>lo'`` / ``'<axis>hi'``.
    """

    def __init__(self, n, lo, hi, periodic=True, warp=0.0, curved=0.0):
        self.n = n = tuple(int(v) for v in n)
        self.ndims = nd = len(n)
        self.etype = 'quad' if nd == 2 else 'hex'
        self.lo = np.broadcast_to(np.asarray(lo, dtype=float), (nd,))
        self.hi = np.broadcast_to(np.asarray(hi, dtype=float), (nd,))
        self.periodic = ((periodic,)*nd if isinstance(periodic, bool)
                         else tuple(periodic))
        self.warp = warp
        # Fraction of each partition's elements flagged as curved (they
        # then take the stored-metric kernel path; geometrically they are
        # the same multilinear cells)
        self.curved = curved
        self.neles = int(np.prod(n))

        nfaces = 2*nd
        bnames = [f'{"xyz"[a]}{s}' for a in range(nd) if not self.periodic[a]
                  for s in ('lo', 'hi')]
        self.codec = ([f'eles/{self.etype}'] +
                      [f'eles/{self.etype}/face/{f}' for f in range(nfaces)] +
                      [f'bc/{b}' for b in bnames])
        self.cidxmap = {f + 1: (self.etype, f) for f in range(nfaces)}

        # Neighbour table: for element g, face f -> (codec idx, global ele)
        ijk = np.indices(n[::-1])[::-1].reshape(nd, -1)   # ijk[0] fastest
        self.ijk = ijk
        self.rcidx = np.empty((self.neles, nfaces), dtype=np.int16)
        self.roff = np.empty((self.neles, nfaces), dtype=np.int64)
        strides = np.cumprod((1,) + n[:-1])

        for f, (ax, sgn) in enumerate(_face_dirs[self.etype]):
            c = ijk.copy()
            c[ax] += sgn
            out = (c[ax] < 0) | (c[ax] >= n[ax])
            c[ax] %= n[ax]
            self.roff[:, f] = strides @ c
            self.rcidx[:, f] = _face_opp[self.etype][f] + 1

            if not self.periodic[ax]:
                bn = f'bc/{"xyz"[ax]}{"lo" if sgn < 0 else "hi"}'
                self.rcidx[out, f] = self.codec.index(bn)
                self.roff[out, f] = -1

    def vertices(self, g):
        """Vertex coordinates ``(2^nd, len(g), nd)`` of elements ``g`` in
        reference vertex order (first coordinate fastest)."""
        nd, n = self.ndims, self.n
        h = (self.hi - self.lo)/np.asarray(n)
        ijk = self.ijk[:, g]

        corners = np.indices((2,)*nd)[::-1].reshape(nd, -1)   # (nd, 2^nd)
        idx = ijk[:, None, :] + corners[:, :, None]           # (nd, 2^nd, m)
        x = self.lo[:, None, None] + idx*h[:, None, None]

        if self.warp:
            # Smooth, box-periodic displacement -> non-affine linear cells
            # (phase offsets keep the vertices of coarse meshes -- two or
            # three cells per direction -- away from the zeros of the sines)
            L = (self.hi - self.lo)[:, None, None]
            ph = 2*np.pi*(x - self.lo[:, None, None])/L
            ph = ph + np.array([0.5, 0.9, 1.3][:nd])[:, None, None]
            s = np.prod(np.sin(ph), axis=0)
            x = x + self.warp*h[:, None, None]*s*np.array(
                [1.0, -0.7, 0.5][:nd])[:, None, None]

        return np.ascontiguousarray(x.transpose(1, 2, 0))

    def brick_partition(self, parts):
        """Assign elements to ``prod(parts)`` equal bricks."""
        parts = tuple(parts)
        pid = np.zeros(self.neles, dtype=np.int32)
        mul = 1

        for ax, p in enumerate(parts):
            if self.n[ax] % p:
                raise ValueError('Brick partition must divide the box')
            pid += mul*(self.ijk[ax] // (self.n[ax] // p))
            mul *= p

        return pid

    def partition_order(self, vparts):
        """Global element numbers of each partition in storage order."""
        vparts = np.asarray(vparts)
        nparts = int(vparts.max()) + 1

        internal = np.ones(self.neles, dtype=bool)
        for f in range(self.roff.shape[1]):
            nb = self.roff[:, f]
            cut = (nb >= 0) & (vparts[np.maximum(nb, 0)] != vparts)
            internal[cut] = False

        order = np.lexsort((internal, vparts))
        bounds = np.searchsorted(vparts[order], np.arange(nparts + 1))

        return [order[bounds[p]:bounds[p + 1]] for p in range(nparts)]

    def local_mesh(self, vparts=None, rank=0):
        """The ``Mesh`` partition ``rank`` sees."""
        if vparts is None:
            vparts = np.zeros(self.neles, dtype=np.int32)

        vparts = np.asarray(vparts)
        et = self.etype
        gidx = self.partition_order(vparts)[rank]
        nloc, nfaces = len(gidx), self.roff.shape[1]

        mesh = Mesh(ndims=self.ndims, codec=self.codec, etypes=[et],
                    cidxmap=self.cidxmap)
        mesh.eidxs[et] = gidx
        mesh.spts[et] = self.vertices(gidx)
        mesh.spts_curved[et] = np.arange(nloc) < int(round(self.curved*nloc))

        # Global -> local numbering for this partition
        g2l = np.full(self.neles, -1, dtype=np.int64)
        g2l[gidx] = np.arange(nloc)

        # Flatten faces: face-major, then local element
        lcidx = np.repeat(np.arange(1, nfaces + 1, dtype=np.int16), nloc)
        leidx = np.tile(np.arange(nloc), nfaces)
        lgidx = np.tile(gidx, nfaces)
        rcidx = self.rcidx[gidx].T.ravel()
        rgidx = self.roff[gidx].T.ravel()
        reidx = np.where(rgidx >= 0, g2l[np.maximum(rgidx, 0)], -1)

        is_bnd = rgidx == -1
        is_loc = reidx >= 0
        is_mpi = ~(is_bnd | is_loc)

        # Interior faces, listed once
        stride = max(leidx[is_loc].max(initial=-1),
                     reidx[is_loc].max(initial=-1)) + 1
        lkey = lcidx[is_loc].astype(np.int64)*stride + leidx[is_loc]
        rkey = rcidx[is_loc].astype(np.int64)*stride + reidx[is_loc]
        keep = np.flatnonzero(is_loc)[lkey < rkey]

        mesh.con = (Connectivity(lcidx[keep], leidx[keep], self.cidxmap),
                    Connectivity(rcidx[keep], reidx[keep], self.cidxmap))

        # Boundary faces
        for bc in np.unique(rcidx[is_bnd]):
            sel = is_bnd & (rcidx == bc)
            mesh.bcon[self.codec[bc][3:]] = Connectivity(
                lcidx[sel], leidx[sel], self.cidxmap
            )

        # Inter-partition faces
        if is_mpi.any():
            m = np.flatnonzero(is_mpi)
            nbr = vparts[rgidx[m]]
            gstride = self.neles

            for p in np.unique(nbr):
                ix = m[nbr == p]

                if rank < p:
                    key = rcidx[ix].astype(np.int64)*gstride + rgidx[ix]
                else:
                    key = lcidx[ix].astype(np.int64)*gstride + lgidx[ix]

                ix = ix[np.argsort(key, kind='stable')]
                mesh.con_p[int(p)] = Connectivity(lcidx[ix], leidx[ix],
                                                  self.cidxmap)

        return mesh


# -- mixed element boxes (BASELINE.json configs[3]) ---------------------------
# Vertex lists in std-element order (pyfr/shapes.py std_ele(1)) as corner
# offsets of the unit cell; 'P' is the cell centre.  Every element has a
# positive Jacobian.
_TRI_PAIR = [[(0, 0), (1, 0), (0, 1)], [(1, 1), (0, 1), (1, 0)]]

# Quadrilateral bases of the six pyramids of a cell, wound so that the
# normal of (V1 - V0) x (V2 - V0) points at the cell centre
_PYR_BASES = [
    [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0)],       # z lo
    [(0, 0, 1), (0, 1, 1), (1, 0, 1), (1, 1, 1)],       # z hi
    [(0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1)],       # x lo
    [(1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)],       # x hi
    [(0, 0, 0), (0, 0, 1), (1, 0, 0), (1, 0, 1)],       # y lo
    [(0, 1, 0), (1, 1, 0), (0, 1, 1), (1, 1, 1)],       # y hi
]


def _cell_elements(kind):
    if kind == 'quad':
        return {'quad': [[(0, 0), (1, 0), (0, 1), (1, 1)]]}
    if kind == 'tri':
        return {'tri': _TRI_PAIR}
    if kind == 'hex':
        return {'hex': [[(i, j, k) for k in (0, 1) for j in (0, 1)
                         for i in (0, 1)]]}
    if kind == 'pri':
        return {'pri': [[v + (k,) for k in (0, 1) for v in t]
                        for t in _TRI_PAIR]}
    if kind == 'pyr':
        return {'pyr': [b + ['P'] for b in _PYR_BASES]}
    if kind == 'pyt':
        # top and bottom pyramids split into two tetrahedra on the V0-V3
        # diagonal of their bases (the same physical diagonal on both)
        tets = []
        for v0, v1, v2, v3 in _PYR_BASES[:2]:
            tets += [[v0, v1, v3, 'P'], [v0, v3, v2, 'P']]
        return {'pyr': [b + ['P'] for b in _PYR_BASES[2:]], 'tet': tets}
    raise ValueError(f'Unknown cell kind {kind!r}')


class MixedBoxMesh:
    """Fully periodic conforming box of several element types.

    ``kinds[i, j(, k)]`` is the kind of each unit cell: ``quad`` | ``tri``
    (cell cut on its anti-diagonal) in 2-D; ``hex`` | ``pri`` (that triangle
    pair extruded) | ``pyr`` (six pyramids about the cell centre) | ``pyt``
    (ditto, top and bottom pyramid cut into two tetrahedra each) in 3-D.
    Cells stacked in z must share their kind (``columns`` builds such
    arrays) so that neighbouring columns meet in whole quadrilaterals.
    Partitions are bricks of cells (``brick_partition``)."""

    def __init__(self, kinds, h=1.0, warp=0.0):
        self.kinds = kinds = np.asarray(kinds, dtype=object)
        self.ndims = nd = kinds.ndim
        self.n, self.h, self.warp = kinds.shape, h, warp
        self.L = np.array(self.n, dtype=float)*h

        verts, cells = {}, {}
        for idx in np.ndindex(*self.n):
            org = np.array(idx, dtype=float)
            for et, lists in _cell_elements(kinds[idx]).items():
                for vl in lists:
                    verts.setdefault(et, []).append(
                        [(org + 0.5 if isinstance(v, str) else
                          org + np.array(v))*h for v in vl]
                    )
                    cells.setdefault(et, []).append(idx)

        # Cell (i, j[, k]) of every element
        self._cell = {et: np.array(c) for et, c in cells.items()}

        self.etypes = etypes = sorted(verts)
        # (nverts, neles, ndims), unwarped: face pairing works on these
        self._x0 = {et: np.array(verts[et]).swapaxes(0, 1) for et in etypes}

        fv = {et: _face_verts(et) for et in etypes}
        self.codec = [f'eles/{et}' for et in etypes]
        for et in etypes:
            self.codec += [f'eles/{et}/face/{f}' for f in range(len(fv[et]))]
        self.cidxmap = {self.codec.index(f'eles/{et}/face/{f}'): (et, f)
                        for et in etypes for f in range(len(fv[et]))}

        # Pair faces through their centroids (vertex means), folded into
        # the period and scaled so that every centroid is an integer
        table = {}
        per = 12*np.array(self.n)
        for et in etypes:
            for f, ids in enumerate(fv[et]):
                cen = self._x0[et][ids].mean(axis=0)
                keys = np.rint(cen*12/h).astype(np.int64) % per
                for e, k in enumerate(map(tuple, keys)):
                    table.setdefault(k, []).append((et, e, f))

        if any(len(v) != 2 for v in table.values()):
            raise ValueError('Cell kinds do not form a conforming mesh')

        self.faces = {et: np.zeros((self._x0[et].shape[1], len(fv[et]), 2),
                                   dtype=np.int64) for et in etypes}
        for a, b in table.values():
            for (et, e, f), (net, ne, nf) in ((a, b), (b, a)):
                self.faces[et][e, f] = (
                    self.codec.index(f'eles/{net}/face/{nf}'), ne
                )

    @staticmethod
    def columns(nx, ny, nz, pattern):
        """Kinds for an ``nx x ny (x nz)`` box: column (i, j) gets
        ``pattern[(i + 2 j) % len(pattern)]``."""
        k2 = np.array([[pattern[(i + 2*j) % len(pattern)] for j in range(ny)]
                       for i in range(nx)], dtype=object)
        return k2 if nz is None else np.repeat(k2[:, :, None], nz, axis=2)

    def vertices(self, et):
        x = self._x0[et]

        if self.warp:
            nd = self.ndims
            ph = 2*np.pi*x/self.L
            x = x + self.warp*self.h*np.stack(
                [np.sin(ph[..., (a + 1) % nd] + 0.5 + 0.4*a)
                 for a in range(nd)], axis=-1
            )

        return np.ascontiguousarray(x)

    def brick_partition(self, parts):
        """``{etype: partition of each element}`` for ``prod(parts)`` equal
        bricks of cells."""
        parts = tuple(parts)
        if any(n % p for n, p in zip(self.n, parts)):
            raise ValueError('Brick partition must divide the box')

        out = {}
        for et, c in self._cell.items():
            pid, mul = np.zeros(len(c), dtype=np.int32), 1
            for ax, p in enumerate(parts):
                pid += mul*(c[:, ax] // (self.n[ax] // p))
                mul *= p
            out[et] = pid

        return out

    def partition_order(self, vparts):
        """Per rank and element type, the global element numbers in storage
        order: partition-boundary elements first, then by number
        (pyfr/partitioners/base.py:286-290)."""
        nparts = max(int(v.max()) for v in vparts.values()) + 1
        etidx = {self.codec.index(f'eles/{et}/face/{f}'): et
                 for et in self.etypes
                 for f in range(self.faces[et].shape[1])}

        order = [{} for _ in range(nparts)]
        for et in self.etypes:
            fc = self.faces[et]
            nbp = np.empty(fc.shape[:2], dtype=np.int32)
            for f in range(fc.shape[1]):
                for c in np.unique(fc[:, f, 0]):
                    sel = fc[:, f, 0] == c
                    nbp[sel, f] = vparts[etidx[c]][fc[sel, f, 1]]

            internal = np.all(nbp == vparts[et][:, None], axis=1)
            idx = np.lexsort((internal, vparts[et]))
            bounds = np.searchsorted(vparts[et][idx], np.arange(nparts + 1))
            for p in range(nparts):
                if bounds[p + 1] > bounds[p]:
                    order[p][et] = idx[bounds[p]:bounds[p + 1]]

        return order

    def local_mesh(self, vparts=None, rank=0):
        if vparts is None:
            vparts = {et: np.zeros(self._x0[et].shape[1], dtype=np.int32)
                      for et in self.etypes}

        gidx = self.partition_order(vparts)[rank]
        etypes = [et for et in self.etypes if et in gidx]
        etof = {c: et for c, (et, f) in self.cidxmap.items()}

        mesh = Mesh(ndims=self.ndims, codec=self.codec, etypes=etypes,
                    cidxmap=self.cidxmap, uuid='mixed')

        g2l = {}
        for et in etypes:
            g = gidx[et]
            mesh.eidxs[et] = g
            mesh.spts[et] = self.vertices(et)[:, g]
            mesh.spts_curved[et] = np.zeros(len(g), dtype=bool)
            g2l[et] = np.full(self._x0[et].shape[1], -1, dtype=np.int64)
            g2l[et][g] = np.arange(len(g))

        # Flatten: type by type, face by face, element by element
        # (pyfr/readers/native.py:430-443)
        parts = []
        for et in etypes:
            g, ne = gidx[et], len(gidx[et])
            for f in range(self.faces[et].shape[1]):
                lc = self.codec.index(f'eles/{et}/face/{f}')
                parts.append((np.full(ne, lc), np.arange(ne), g,
                              self.faces[et][g, f, 0],
                              self.faces[et][g, f, 1]))

        lcidx, leidx, lgidx, rcidx, rgidx = map(np.concatenate, zip(*parts))

        # Neighbour's partition and (where it is ours) local number
        rpart = np.empty(len(rcidx), dtype=np.int32)
        reidx = np.full(len(rcidx), -1, dtype=np.int64)
        for c in np.unique(rcidx):
            sel, net = rcidx == c, etof[c]
            rpart[sel] = vparts[net][rgidx[sel]]
            if net in g2l:
                reidx[sel] = g2l[net][rgidx[sel]]

        is_loc = reidx >= 0

        # Interior faces once, smaller (cidx, element) key on the left
        stride = max(leidx[is_loc].max(initial=-1),
                     reidx[is_loc].max(initial=-1)) + 1
        lkey = lcidx[is_loc]*stride + leidx[is_loc]
        rkey = rcidx[is_loc]*stride + reidx[is_loc]
        keep = np.flatnonzero(is_loc)[lkey < rkey]

        mesh.con = (Connectivity(lcidx[keep], leidx[keep], self.cidxmap),
                    Connectivity(rcidx[keep], reidx[keep], self.cidxmap))

        # Inter-partition faces: both sides sort on the lower rank's
        # (cidx, global element) key (pyfr/readers/native.py:518-521)
        m = np.flatnonzero(~is_loc)
        gstride = max(x.shape[1] for x in self._x0.values())
        for p in np.unique(rpart[m]):
            ix = m[rpart[m] == p]

            if rank < p:
                key = rcidx[ix]*gstride + rgidx[ix]
            else:
                key = lcidx[ix]*gstride + lgidx[ix]

            ix = ix[np.argsort(key, kind='stable')]
            mesh.con_p[int(p)] = Connectivity(lcidx[ix], leidx[ix],
                                              self.cidxmap)

        return mesh


def _face_verts(etype):
    """Vertex numbers (std-element order) on each face of a linear
    element."""
    shape = shape_map[etype]

    if hasattr(shape, 'faces') and shape.faces is not None:
        corners = {'line': [(-1,), (1,)],
                   'quad': [(-1, -1), (1, -1), (-1, 1), (1, 1)]}
        lin = np.asarray(shape.std_ele(1), dtype=float)
        return [[int(np.argmin(np.abs(lin - np.array(proj(*c),
                                                     dtype=float)).sum(1)))
                 for c in corners[kind]]
                for kind, proj, _ in shape.faces]
    else:
        return shape(_nverts[etype], _LinCfg(etype)).faceverts


_nverts = {'tri': 3, 'tet': 4, 'pri': 6, 'pyr': 5}


class _LinCfg:
    """Just enough configuration to look a tabulated shape up."""

    def __init__(self, etype):
        from pyfr_b200.host.shapes import TabulatedShape
        self._rule = TabulatedShape._rules[etype]

    def getint(self, sect, opt, default=None):
        return 1

    def get(self, sect, opt, default=None):
        return self._rule if opt == 'soln-pts' else default
