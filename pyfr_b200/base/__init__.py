"""Backend-base contract shared by the B200 backend and the test oracle.

By default this is the in-tree mirror of ``pyfr.backends.base`` (same
names, so host code reads identically against either package).  With
``PYFR_B200_BASE=pyfr.backends.base`` in the environment the *reference's
own* base classes are used instead: that is the configuration in which the
backend is a drop-in inside a PyFR checkout, and the one
``tests/test_reference_dropin.py`` exercises where /root/reference exists.
"""

import os

_src = os.environ.get('PYFR_B200_BASE', '')

if _src:
    from importlib import import_module

    _b = import_module(f'{_src}.backend')
    _p = import_module(f'{_src}.provider')
    _t = import_module(f'{_src}.types')

    BaseBackend, MemoryInfo = _b.BaseBackend, _b.MemoryInfo
    Kernel, NullKernel = _p.Kernel, _p.NullKernel
    NotSuitableError = _p.NotSuitableError
    MetaKernel = _p.BaseMetaKernel
    OrderedMetaKernel = _p.BaseOrderedMetaKernel
    UnorderedMetaKernel = _p.BaseUnorderedMetaKernel
    Graph = _t.Graph
    ConstMatrix, Extent, Matrix = _t.ConstMatrix, _t.Extent, _t.Matrix
    MatrixBase, MatrixSlice = _t.MatrixBase, _t.MatrixSlice
    StorageRegion, View = _t.StorageRegion, _t.View
    XchgMatrix, XchgView = _t.XchgMatrix, _t.XchgView
else:
    from pyfr_b200.base.backend import BaseBackend, MemoryInfo
    from pyfr_b200.base.kernels import (Graph, Kernel, MetaKernel,
                                        NotSuitableError, NullKernel,
                                        OrderedMetaKernel,
                                        UnorderedMetaKernel)
    from pyfr_b200.base.types import (ConstMatrix, Extent, Matrix,
                                      MatrixBase, MatrixSlice, StorageRegion,
                                      View, XchgMatrix, XchgView)
