"""CUDA source for the register-bank update ``x0 = a0*x0 + a1*x1 + ...``.

Counterpart of ``pyfr/backends/cuda/kernels/axnpby.mako:4-44``: the same
``a0 == 0`` (overwrite, ``x0`` never read) and general paths, but written as
a flat grid-stride sweep over the whole allocation -- every bank shares one
blocked layout, so no index arithmetic is needed -- with coalesced
element-wide accesses, four independent iterations of a thread in flight
(``#pragma unroll``); an HBM-streaming kernel, not on the RHS path.
"""

from pyfr_b200.kernels import physics as ph

AXNPBY_VEC = 4


def axnpby_source(be, nv):
    xs = ', '.join(f'const fpdtype_t* __restrict__ x{i}' for i in range(1, nv))
    as_ = ', '.join(f'fpdtype_t a{i}' for i in range(nv))
    rest = ' + '.join(f'a{i}*x{i}[j]' for i in range(1, nv)) or 'FP(0.0)'

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz)}

extern "C" __global__ void __launch_bounds__(256)
axnpby(long long n, fpdtype_t* __restrict__ x0{', ' + xs if xs else ''}, {as_})
{{
    const long long stride = (long long) gridDim.x*blockDim.x;
    long long j = (long long) blockIdx.x*blockDim.x + threadIdx.x;

    if (a0 == FP(0.0))
    {{
        #pragma unroll {AXNPBY_VEC}
        for (; j < n; j += stride)
            x0[j] = {rest};
    }}
    else
    {{
        #pragma unroll {AXNPBY_VEC}
        for (; j < n; j += stride)
            x0[j] = a0*x0[j] + {rest};
    }}
}}
'''
    return src, 'axnpby'
