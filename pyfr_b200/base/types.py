"""Backend data types: extents, matrices, slices, views, exchange buffers.

This is a from-scratch mirror of the *contract* of the reference's
``pyfr/backends/base/types.py`` (Extent :19-37, MatrixBase :40-150,
MatrixSlice :169-217, StorageRegion :220-238, ConstMatrix :241-245,
XchgMatrix :248-257, View :260-320, XchgView :323-340).  Host code written
against the reference API (``backend.matrix(...)``, ``m.get()``,
``m.slice(...)``, ``backend.view(...)``) works unchanged; the storage
layout rules are bit-for-bit those of the reference so that view
``mapping``/``rstrides`` index arrays are identical.

Layout rules (``k = soasz``, ``c = csubsz``):

* stacked matrix ``ioshape = (..., nvar, narr)``: ``narr`` is padded to a
  multiple of ``c``.  Unblocked: row-major ``[nrow][narr/k][nvar][k]`` with
  ``leaddim = nvar*narr_padded``.  Blocked: ``[narr/c][nrow][c/k][nvar][k]``
  with ``leaddim = nvar*c`` and ``blocksz = nrow*leaddim``.
* plain 2-D matrix: unblocked ``[nrow][ncol]`` (``ncol`` padded to ``c`` if
  tagged ``align``); blocked ``[ncol/c][nrow][c]``.  Tags ``xchg`` and
  ``noblock`` force the unblocked form.
"""

import numpy as np


def _roundup(n, m):
    return -(-n // m)*m


class _Storage:
    @property
    def storage_root(self):
        return self._storage_root

    def same_storage(self, other):
        return self.storage_root is other.storage_root


class Extent(_Storage):
    """A named allocation which several matrices are carved out of."""

    def __init__(self, name=None):
        self.name = name
        self.offset = 0
        self.nbytes = 0
        self.basedata = None
        self._clients = []
        self._storage_root = self

    def reserve(self, obj, nbytes):
        self._clients.append((obj, self.nbytes))
        self.nbytes += nbytes

    def commit(self, alloc_fn):
        self.basedata = alloc_fn(self.nbytes)

        for obj, off in self._clients:
            obj.onalloc(self.basedata, off)
            obj._storage_root = self

        self._clients = []


class MatrixBase(_Storage):
    _base_tags = frozenset()

    def __init__(self, backend, dtype, ioshape, initval, extent, tags):
        self.backend = backend
        self.tags = set(self._base_tags) | set(tags)
        self.dtype = dtype
        self.itemsize = np.dtype(dtype).itemsize
        self.ioshape = ioshape = tuple(ioshape)

        k, c = backend.soasz, backend.csubsz

        if len(ioshape) == 2:
            nrow, ncol = ioshape
            blocked = backend.blocks and not (self.tags & {'xchg', 'noblock'})

            if blocked:
                leaddim = c
            elif 'align' in self.tags:
                leaddim = _roundup(ncol, c)
            else:
                leaddim = ncol

            nblocks = _roundup(ncol, leaddim) // leaddim if leaddim else 0
            datashape = [nblocks, nrow, leaddim]
        elif len(ioshape) in (3, 4):
            nvar, narr = ioshape[-2:]
            nparr = _roundup(narr, c)

            nrow = int(np.prod(ioshape[:-2]))
            ncol = nvar*nparr
            leaddim = nvar*c if backend.blocks else ncol
            nblocks = _roundup(ncol, leaddim) // leaddim if leaddim else 0
            datashape = [nblocks, *ioshape[:-2], nparr // (nblocks*k) if
                         nblocks else 0, nvar, k]
        else:
            raise ValueError('Invalid matrix dimensionality')

        self.nrow, self.ncol, self.leaddim = nrow, ncol, leaddim
        self.nblocks, self.datashape = nblocks, datashape
        self.blocksz = nrow*leaddim
        self.splitsz = leaddim if backend.blocks else k
        self.nbytes = nblocks*self.blocksz*self.itemsize
        self.traits = (nblocks, nrow, ncol, leaddim, dtype)

        if initval is not None:
            if tuple(initval.shape) != ioshape:
                raise ValueError('Invalid initial value')

            self._initval = np.asanyarray(initval, dtype=dtype)
        else:
            self._initval = None

        backend.malloc(self, extent)

    # -- host <-> device element ordering ---------------------------------
    def _pack(self, ary):
        """``ioshape`` array -> flat array in storage order."""
        k, c = self.backend.soasz, self.backend.csubsz
        ary = np.asarray(ary, dtype=self.dtype)

        if ary.ndim == 2:
            pad = self.nblocks*self.leaddim - ary.shape[1]
            a = np.pad(ary, [(0, 0), (0, pad)])
            a = a.reshape(self.nrow, self.nblocks, self.leaddim)
        else:
            nvar, narr = ary.shape[-2:]
            nparr = _roundup(narr, c)
            a = np.pad(ary.reshape(self.nrow, nvar, narr),
                       [(0, 0), (0, 0), (0, nparr - narr)])
            # (nrow, nvar, nparr/k, k) -> (nrow, nparr/k, nvar, k)
            a = a.reshape(self.nrow, nvar, nparr // k, k).transpose(0, 2, 1, 3)
            a = a.reshape(self.nrow, self.nblocks, self.leaddim)

        return np.ascontiguousarray(a.transpose(1, 0, 2)).reshape(-1)

    def _unpack(self, flat):
        """Flat storage-order array -> ``ioshape`` array."""
        k = self.backend.soasz
        a = np.asarray(flat).reshape(self.nblocks, self.nrow, self.leaddim)
        a = a.transpose(1, 0, 2)

        if len(self.ioshape) == 2:
            a = a.reshape(self.nrow, -1)[:, :self.ioshape[1]]
        else:
            nvar, narr = self.ioshape[-2:]
            a = a.reshape(self.nrow, -1, nvar, k).transpose(0, 2, 1, 3)
            a = a.reshape(self.nrow, nvar, -1)[..., :narr]
            a = a.reshape(self.ioshape)

        return np.ascontiguousarray(a)

    def get(self):
        if hasattr(self, '_initval'):
            if self._initval is not None:
                return self._initval
            else:
                return np.zeros(self.ioshape, dtype=self.dtype)
        else:
            return self._get()

    def _get(self):
        raise NotImplementedError

    def slice(self, ra=None, rb=None, ca=None, cb=None):
        ra, rb = ra or 0, rb or self.nrow
        ca, cb = ca or 0, cb or self.ncol

        return self.backend.matrix_slice(self, ra, rb, ca, cb)


class Matrix(MatrixBase):
    def set(self, ary):
        if tuple(ary.shape) != self.ioshape:
            raise ValueError('Invalid matrix shape')

        if hasattr(self, '_initval'):
            self._initval = np.asanyarray(ary, dtype=self.dtype)
        else:
            self._set(ary)

    def _set(self, ary):
        raise NotImplementedError


class ConstMatrix(MatrixBase):
    _base_tags = frozenset({'const'})

    def __init__(self, backend, dtype, initval, tags):
        super().__init__(backend, dtype, initval.shape, initval, None, tags)


class XchgMatrix(Matrix):
    _base_tags = frozenset({'xchg'})

    # The reference hands back persistent MPI requests here
    # (types.py:250-257); backends of this package return their own
    # exchange descriptors, see pyfr_b200.types.B200XchgMatrix.
    def recvreq(self, comm, pid, tag):
        return comm.recv_init(self, pid, tag)

    def sendreq(self, comm, pid, tag):
        return comm.send_init(self, pid, tag)


class MatrixSlice(_Storage):
    def __init__(self, backend, mat, ra, rb, ca, cb):
        if ra < 0 or rb > mat.nrow or rb < ra:
            raise ValueError('Invalid row slice')
        if ca < 0 or cb > mat.ncol or cb < ca:
            raise ValueError('Invalid column slice')
        if ca % mat.splitsz:
            raise ValueError('Starting column must conform to backend '
                             'alignment requirements')

        self.backend, self.parent = backend, mat
        self.ra, self.rb, self.ca, self.cb = int(ra), int(rb), int(ca), int(cb)
        self.nrow, self.ncol = self.rb - self.ra, self.cb - self.ca
        self.dtype, self.itemsize = mat.dtype, mat.itemsize
        self.leaddim, self.blocksz = mat.leaddim, mat.blocksz
        self.nblocks = _roundup(self.ncol, self.leaddim) // self.leaddim
        self.traits = (self.nblocks, self.nrow, self.ncol, self.leaddim,
                       self.dtype)
        self.tags = mat.tags | {'slice'}

        if backend.blocks:
            self.ba, self.bb = self.ca // self.leaddim, self.cb // self.leaddim

        # Full-width slices are contiguous enough to memcpy
        if ca == 0 and cb == mat.ncol:
            self.nbytes = self.nrow*self.leaddim*self.nblocks*self.itemsize

    @property
    def basedata(self):
        return self.parent.basedata

    @property
    def offset(self):
        if self.backend.blocks:
            rel = self.ba*self.blocksz + self.ra*self.leaddim
        else:
            rel = self.ra*self.leaddim + self.ca

        return self.parent.offset + rel*self.itemsize

    @property
    def storage_root(self):
        return self.parent.storage_root


class StorageRegion(_Storage):
    def __init__(self, parent, offset, nbytes):
        offset, nbytes = int(offset), int(nbytes)

        if offset < 0 or nbytes < 0 or offset + nbytes > parent.nbytes:
            raise ValueError('Invalid storage region')

        self.parent, self.rel_offset, self.nbytes = parent, offset, nbytes

    @property
    def basedata(self):
        return self.parent.basedata

    @property
    def offset(self):
        return self.parent.offset + self.rel_offset

    @property
    def storage_root(self):
        return self.parent.storage_root


class View:
    """Gather/scatter index set over matrices sharing one storage root.

    ``mapping[i]`` is the element offset (from the storage root) of view
    point ``i``, variable 0; variable ``v`` is ``soasz*v`` further on and
    view row ``r`` is ``rstrides[i]*r`` further on (reference
    types.py:294-320).
    """

    def __init__(self, backend, matmap, rmap, cmap, rstridemap, vshape, tags):
        matmap, rmap, cmap = map(np.asarray, (matmap, rmap, cmap))

        self.n = len(matmap)
        self.nvrow = vshape[-2] if len(vshape) == 2 else 1
        self.nvcol = vshape[-1] if len(vshape) >= 1 else 1

        self._mats = mats = [backend.mats[i] for i in np.unique(matmap)]
        m0 = mats[0]

        self.storage_root = m0.storage_root
        self.basedata = m0.basedata
        self.refdtype = m0.dtype

        oktypes = (backend.matrix_cls, backend.matrix_slice_cls)
        if not all(isinstance(m, oktypes) for m in mats):
            raise TypeError('Incompatible matrix type for view')
        if not all(m.same_storage(m0) for m in mats):
            raise TypeError('All viewed matrices must belong to the same '
                            'storage object')
        if not all(m.dtype == m0.dtype for m in mats):
            raise TypeError('Mixed data types are not supported')

        ixdtype = backend.ixdtype
        k, c = backend.soasz, backend.csubsz

        base = np.empty(self.n, dtype=ixdtype)
        ld = np.empty(self.n, dtype=ixdtype)

        for m in mats:
            sel = matmap == m.mid
            cm = cmap[sel]
            base[sel] = (m.offset // m.itemsize
                         + (cm*self.nvcol // m.leaddim)*m.blocksz)
            ld[sel] = m.leaddim

        cm = cmap % c if backend.blocks else cmap
        mapping = base + rmap*ld + (cm // k)*(k*self.nvcol) + cm % k

        self.mapping = backend.const_matrix(mapping[None, :].astype(ixdtype),
                                            dtype=ixdtype, tags=tags)

        if self.nvrow > 1:
            rstrides = (rstridemap*ld)[None, :].astype(ixdtype)
            self.rstrides_val = int(rstrides.flat[0]) if self.n else 0
            self.rstrides = backend.const_matrix(rstrides, dtype=ixdtype,
                                                 tags=tags)
        else:
            self.rstrides_val = 0
            self.rstrides = None


class XchgView:
    def __init__(self, backend, matmap, rmap, cmap, rstridemap, vshape, tags):
        self.view = v = backend.view(matmap, rmap, cmap, rstridemap, vshape,
                                     tags)
        self.n, self.nvrow, self.nvcol = v.n, v.nvrow, v.nvcol
        self.xchgmat = backend.xchg_matrix((v.nvrow, v.nvcol*v.n), tags=tags)

    def recvreq(self, comm, pid, tag):
        return self.xchgmat.recvreq(comm, pid, tag)

    def sendreq(self, comm, pid, tag):
        return self.xchgmat.sendreq(comm, pid, tag)
