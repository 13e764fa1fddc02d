"""Backend-base contract shared by the B200 backend and the test oracle.

Same names as ``pyfr.backends.base`` in the reference so that host code
reads identically against either package.
"""

from pyfr_b200.base.backend import BaseBackend, MemoryInfo
from pyfr_b200.base.kernels import (Graph, Kernel, MetaKernel,
                                    NotSuitableError, NullKernel,
                                    OrderedMetaKernel, UnorderedMetaKernel)
from pyfr_b200.base.types import (ConstMatrix, Extent, Matrix, MatrixBase,
                                  MatrixSlice, StorageRegion, View, XchgMatrix,
                                  XchgView)
