"""Backend base class: allocation bookkeeping and kernel dispatch.

Mirror of the contract of ``pyfr/backends/base/backend.py:32-225`` in the
reference: configuration keys (``[backend] precision, memory-model``),
extent-based allocation (``malloc``/``commit``), matrix factories with
matrix ids recorded in ``self.mats``, const-matrix de-duplication and
provider-based ``kernel()`` dispatch with ``NotSuitableError`` fall-through.
"""

from collections import namedtuple
from contextlib import contextmanager
from itertools import count
from weakref import WeakSet, WeakValueDictionary

import numpy as np

from pyfr_b200.base.kernels import NotSuitableError
from pyfr_b200.base.types import Extent, StorageRegion


MemoryInfo = namedtuple('MemoryInfo', ['current', 'peak', 'free', 'total'])


class BaseBackend:
    name = None
    has_double = True

    # Slots concrete backends fill in
    const_matrix_cls = matrix_cls = matrix_slice_cls = None
    view_cls = xchg_matrix_cls = xchg_view_cls = graph_cls = None
    ordered_meta_kernel_cls = unordered_meta_kernel_cls = None

    def __init__(self, cfg):
        self.cfg = cfg

        prec = cfg.get('backend', 'precision', 'double')
        if prec not in {'single', 'double'}:
            raise ValueError('Backend precision must be either single or '
                             'double')

        self.fpdtype = np.dtype(prec).type
        self.fpdtype_eps = float(np.finfo(self.fpdtype).eps)
        self.fpdtype_max = float(np.finfo(self.fpdtype).max)

        mm = cfg.get('backend', 'memory-model', 'normal')
        if mm == 'normal':
            self.ixdtype = np.int32
        elif mm == 'large':
            self.ixdtype = np.int64
        else:
            raise ValueError('Backend memory model must be either normal '
                             'or large')

        self.autotune_ifac = cfg.getfloat('backend', 'autotune-ifac', 0.95)

        self.mats = WeakValueDictionary()
        self._mids = count()
        self._open_extents = {}
        self._live_extents = WeakSet()
        self._mem_peak = 0
        self._providers = []

    # -- allocation -------------------------------------------------------
    def _padded(self, nbytes):
        return -(-nbytes // self.alignb)*self.alignb

    def malloc(self, obj, extent):
        if extent is None:
            ext = Extent()
            ext.reserve(obj, self._padded(obj.nbytes))
            ext.commit(self._malloc_checked)
            self._note_extent(ext)
        elif isinstance(extent, str):
            ext = self._open_extents.setdefault(extent, Extent(extent))
            ext.reserve(obj, self._padded(obj.nbytes))
        else:
            obj.onalloc(extent.basedata, extent.offset)
            obj._storage_root = extent.storage_root

    def commit(self):
        for ext in self._open_extents.values():
            ext.commit(self._malloc_checked)
            self._note_extent(ext)

        self._open_extents.clear()

    def _malloc_checked(self, nbytes):
        if self.ixdtype == np.int32 and nbytes > 4*2**31 - 1:
            raise RuntimeError('Allocation too large for normal backend '
                               'memory-model')

        return self._malloc_impl(nbytes)

    def _malloc_impl(self, nbytes):
        raise NotImplementedError

    def _note_extent(self, ext):
        self._live_extents.add(ext)
        self._mem_peak = max(self._mem_peak, self._mem_now())

    def _mem_now(self):
        return sum(e.nbytes for e in self._live_extents)

    def memory_info(self):
        return MemoryInfo(self._mem_now(), self._mem_peak, None, None)

    # -- factories --------------------------------------------------------
    def _record(self, m):
        if not hasattr(m, 'mid'):
            m.mid = next(self._mids)
            self.mats[m.mid] = m

        return m

    def const_matrix(self, initval, dtype=None, tags=set()):
        dtype = dtype or self.fpdtype
        initval = np.asanyarray(initval)

        for m in list(self.mats.values()):
            if (isinstance(m, self.const_matrix_cls) and m.dtype == dtype and
                m.ioshape == initval.shape and set(tags) <= m.tags and
                np.array_equal(m.get(), initval)):
                return m

        return self._record(self.const_matrix_cls(self, dtype, initval, tags))

    def matrix(self, ioshape, initval=None, extent=None, tags=set(),
               dtype=None):
        return self._record(self.matrix_cls(self, dtype or self.fpdtype,
                                            ioshape, initval, extent, tags))

    def matrix_slice(self, mat, ra, rb, ca, cb):
        return self._record(self.matrix_slice_cls(self, mat, ra, rb, ca, cb))

    def storage_view(self, parent, offset, nbytes):
        return StorageRegion(parent, offset, nbytes)

    def xchg_matrix(self, ioshape, initval=None, extent=None, tags=set()):
        return self._record(self.xchg_matrix_cls(self, self.fpdtype, ioshape,
                                                 initval, extent, tags))

    def xchg_matrix_for_view(self, view, tags=set()):
        return self.xchg_matrix((view.nvrow, view.nvcol*view.n), tags=tags)

    def view(self, matmap, rmap, cmap, rstridemap=1, vshape=(), tags=set()):
        return self.view_cls(self, matmap, rmap, cmap, rstridemap, vshape,
                             tags)

    def xchg_view(self, matmap, rmap, cmap, rstridemap=1, vshape=(),
                  tags=set()):
        return self.xchg_view_cls(self, matmap, rmap, cmap, rstridemap,
                                  vshape, tags)

    # -- kernels ----------------------------------------------------------
    @contextmanager
    def region(self, name):
        yield

    def kernel(self, name, *args, **kwargs):
        best = None

        for prov in self._providers:
            meth = getattr(prov, name, None)
            if meth is None:
                continue

            try:
                kern = meth(*args, **kwargs)
            except NotSuitableError:
                continue

            if best is None or kern.dt < self.autotune_ifac*best.dt:
                best = kern

                if np.isnan(best.dt):
                    return best

        if best is None:
            raise KeyError(f'Kernel {name!r} has no providers')

        return best

    def ordered_meta_kernel(self, kerns):
        return self.ordered_meta_kernel_cls(kerns)

    def unordered_meta_kernel(self, kerns, splits=None):
        return self.unordered_meta_kernel_cls(kerns, splits)

    def graph(self):
        return self.graph_cls(self)

    def run_kernels(self, kernels, wait=False):
        raise NotImplementedError

    def run_graph(self, graph, wait=False):
        raise NotImplementedError

    def wait(self):
        raise NotImplementedError
