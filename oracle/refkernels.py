"""TEST ORACLE -- the reference's own pointwise kernels, executed.

``make_refkernel_backend`` returns the NumPy oracle backend with its
restated pointwise kernels swapped for the *reference's*: every kernel
template is rendered by ``oracle/minimako.py``, turned into a complete C
kernel by the reference's OpenMP kernel generator
(``pyfr/backends/openmp/generator.py``: argument dereferencing for stacked
matrices, views, broadcasts and 'mpi' arrays included), compiled with gcc
and called block by block through the reference's own argument marshalling
(``BasePointwiseKernelProvider._build_arglst``, ``OpenMPKernelFunction``).
Matrix products, packing and axnpby stay NumPy (the reference uses libxsmm
there, which is plain ``alpha A B + beta C``).

Driven by the reference's host code (``oracle/refharness.py``) this is the
reference itself, minus libxsmm and its C kernel runner: the restated
oracle is compared against it end to end in
tests/test_reference_kernels.py.  Only usable where /root/reference
exists; never imported by product code.
"""

import ctypes as ct
import hashlib
import os
import subprocess
import tempfile

import numpy as np

from oracle import refharness as rh
from oracle.minimako import Renderer
from oracle.npbackend import make_backend

_HEADER = '''
#include <stdint.h>
#include <stdlib.h>
#include <tgmath.h>

#define SOA_SZ {soasz}
#define BLK_SZ {csubsz}

#define min(a, b) ((a) < (b) ? (a) : (b))
#define max(a, b) ((a) > (b) ? (a) : (b))

typedef {fp} fpdtype_t;
typedef {ix} ixdtype_t;

// single-threaded here: the atomics of the reference's header reduce to
#define atomic_min_fpdtype(addr, val) if ((val) < *(addr)) {{ *(addr) = (val); }}
#define atomic_max_fpdtype(addr, val) if ((val) > *(addr)) {{ *(addr) = (val); }}
#define atomic_sum_fpdtype(addr, val) *(addr) += (val)

#define PYFR_FP_PRECISE_BEGIN
'''

_libdir = tempfile.mkdtemp(prefix='pyfr_b200_refk_')
_libs = {}


class _Lib:
    def __init__(self, src):
        key = hashlib.sha256(src.encode()).hexdigest()[:20]
        c, so = (os.path.join(_libdir, f'{key}.{e}') for e in ('c', 'so'))
        with open(c, 'w') as f:
            f.write(src)
        res = subprocess.run(['gcc', '-std=gnu11', '-O1', '-ffp-contract=off',
                              '-w', '-shared', '-fPIC', '-o', so, c, '-lm'],
                             capture_output=True, text=True)
        if res.returncode:
            raise RuntimeError(res.stderr[:4000])
        self.lib = ct.CDLL(so)

    def function(self, name, restype=None, argtypes=None):
        fn = getattr(self.lib, name)
        fn.restype = restype
        return fn


def make_refkernel_backend(rbase, name='oracle-refkernels'):
    rh.install_stubs()
    from pyfr.backends.openmp.generator import OpenMPKernelGenerator
    from pyfr.backends.openmp.provider import (OpenMPKernelFunction,
                                               OpenMPPointwiseKernelProvider)
    from pyfr.nputil import npdtype_to_ctype

    NP = make_backend(rbase, name=name)

    class Fn(OpenMPKernelFunction):
        def __init__(self, backend, fun, argcls, argidxs={}):
            n = len(argcls._fields_)
            self.fun, self.kargs = fun, argcls()
            self._argidxs = argidxs
            self.argsizes, self.subs_offsets = [None]*n, [0]*n
            self.nblocks = None

        def __call__(self):
            # (the reference's C kernel runner does exactly this loop,
            # spread over OpenMP threads)
            for ib in range(self.nblocks):
                self.fun(ct.c_int64(ib) if self._ix64 else ct.c_int32(ib),
                         ct.byref(self.kargs), ct.c_int(0))

        def run(self):
            self()

    class Provider(OpenMPPointwiseKernelProvider):
        def _render_kernel(self, name, mod, extrns, tplargs):
            be = self.backend
            r = Renderer(dict(tplargs), extrns, OpenMPKernelGenerator,
                         be.fpdtype, be.ixdtype)
            r.include(mod)

            hdr = _HEADER.format(soasz=be.soasz, csubsz=be.csubsz,
                                 fp=npdtype_to_ctype(be.fpdtype),
                                 ix=npdtype_to_ctype(be.ixdtype))
            ndim, argn, argt = r.argspecs[name]
            return hdr + r.sources[name], ndim, argn, argt

        def _build_library(self, src):
            if src not in _libs:
                _libs[src] = _Lib(src)
            return _libs[src]

        def _build_kernel(self, kname, src, argtypes, argnames=[]):
            fun = self._build_library(src).function(kname)
            fn = Fn(self.backend, fun, self._get_arg_cls(tuple(argtypes)),
                    {n: i for i, n in enumerate(argnames)})
            fn._ix64 = self.backend.ixdtype == np.int64
            return fn

        def _instantiate_kernel(self, dims, fun, arglst, argm, argv):
            # device pointers of this backend are addresses of NumPy buffers
            def addr(k):
                if isinstance(k, np.ndarray):
                    return k.ctypes.data
                if hasattr(k, 'basedata') and hasattr(k, 'offset'):
                    return k.basedata.ctypes.data + k.offset
                return k

            kern = super()._instantiate_kernel(dims, fun,
                                               [addr(k) for k in arglst],
                                               argm, argv)
            graph_add = lambda graph, deps, k=kern: (
                graph.program.append(('kernel', k)) or k)
            kern.add_to_graph = graph_add
            kern.run = lambda *a, k=kern: k.kernel()
            return kern

    class Backend(NP):
        def __init__(self, cfg):
            super().__init__(cfg)

            # the reference's OpenMP backend always works on blocks
            self.blocks = True
            self.pointwise = Provider(self)
            self._providers = [self._providers[0], self.pointwise]
            self.krunner = None

    Backend.name = name
    return Backend
