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
