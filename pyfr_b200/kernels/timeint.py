"""CUDA source for the kernels either side of the RHS in an explicit
time step: the low-storage Runge-Kutta register update ``rkvdh2``
(``pyfr/integrators/explicit/kernels/rkvdh2.mako``) and the ``reduction``
kernel behind the error estimate and other norms
(``pyfr/backends/cuda/kernels/reduction.mako``,
``pyfr/backends/base/blasext.py:19-46``).

Both are pure streaming kernels over the storage image of register banks
(all banks of an element type share one layout): grid-stride loops over
whole words, so every access is a full 128-byte segment.
"""

import re

from pyfr_b200.kernels import physics as ph


def rkvdh2_source(be, tplargs):
    a, b, e = tplargs['a'], tplargs['b'], tplargs['e']
    stage, nstages = tplargs['stage'], tplargs['nstages']
    errest = tplargs['errest']
    last = stage == nstages - 1

    args = ['long long n', 'fpdtype_t* __restrict__ r1',
            'fpdtype_t* __restrict__ r2']
    if errest:
        args += ['fpdtype_t* __restrict__ rold',
                 'fpdtype_t* __restrict__ rerr']
    args.append('const fpdtype_t* __restrict__ dt_p')

    body = ['const fpdtype_t t1 = r1[i], t2 = r2[i];']
    if errest and stage == 0:
        body += [f'rerr[i] = dt*{ph.fpconst(e[stage])}*t2;', 'rold[i] = t1;']
    elif errest:
        body += [f'rerr[i] = rerr[i] + dt*{ph.fpconst(e[stage])}*t2;']

    if not last:
        body += [f'r1[i] = t1 + dt*{ph.fpconst(a[stage])}*t2;',
                 f'r2[i] = t1 + dt*{ph.fpconst(b[stage])}*t2;']
    else:
        body += [f'r1[i] = t1 + dt*{ph.fpconst(b[stage])}*t2;']

    nl = '\n        '
    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz)}

// stage {stage} of {nstages}{', with error estimate' if errest else ''}
extern "C" __global__ void __launch_bounds__(256)
rkvdh2({', '.join(args)})
{{
    // run-time scalars live in device memory (backend.RuntimeScalars)
    const fpdtype_t dt = *dt_p;
    const long long stride = (long long) gridDim.x*blockDim.x;
    for (long long i = (long long) blockIdx.x*blockDim.x + threadIdx.x;
         i < n; i += stride)
    {{
        {nl.join(body)}
    }}
}}
'''
    return src, 'rkvdh2', [x.split()[-1].lstrip('*') for x in args]


def reduction_source(be, rop, exprs, vnames, svars, pvars, nvars):
    """``exprs`` are C expressions in the variable names ``vnames`` (one
    value per stored word), the scalars ``svars`` and the per-field
    constants ``pvars`` (indexed by the word's field variable)."""
    if rop not in ('sum', 'max'):
        raise ValueError('Invalid reduction operator')

    nex = len(exprs)
    exprs = [re.sub(r'\babs\(', 'fabs(', x) for x in exprs]
    vre = '|'.join(map(re.escape, vnames))
    exprs = [re.sub(rf'\b({vre})\b', r'\1[i]', x) for x in exprs]
    if pvars:
        pre = '|'.join(map(re.escape, pvars))
        exprs = [re.sub(rf'\b({pre})\b', r'c_\1[v]', x) for x in exprs]

    consts = '\n'.join(
        f'__constant__ fpdtype_t c_{k}[{len(v)}] = '
        f'{{{", ".join(ph.fpconst(x) for x in v)}}};'
        for k, v in pvars.items()
    )

    args = (['long long nblocks', 'long long bsz', 'int neles', 'int nrow',
             'int ld'] +
            [f'const fpdtype_t* __restrict__ {v}' for v in vnames] +
            ['fpdtype_t* __restrict__ out'] +
            [f'fpdtype_t {s}' for s in svars])

    init = 'FP(0.0)' if rop == 'sum' else '-FPMAX'
    comb = (lambda a, b: f'{a} + {b}') if rop == 'sum' else \
           (lambda a, b: f'fmax({a}, {b})')
    fpmax = '1.7976931348623157e308' if be.fpdtype.__name__ == 'float64' \
        else '3.4028234e38f'

    acc = '\n        '.join(
        f'acc[{j}] = {comb(f"acc[{j}]", f"({x})")};'
        for j, x in enumerate(exprs)
    )

    if rop == 'sum':
        final = 'atomicAdd(out + j, sdata[0]);'
    elif be.fpdtype.__name__ == 'float64':
        final = '''unsigned long long *p =
                reinterpret_cast<unsigned long long *>(out + j);
            unsigned long long old = *p, cur;
            do
            {
                cur = old;
                if (__longlong_as_double((long long) cur) >= sdata[0])
                    break;
                old = atomicCAS(p, cur, (unsigned long long)
                                __double_as_longlong(sdata[0]));
            } while (old != cur);'''
    else:
        final = '''unsigned int *p = reinterpret_cast<unsigned int *>(out + j);
            unsigned int old = *p, cur;
            do
            {
                cur = old;
                if (__uint_as_float(cur) >= sdata[0])
                    break;
                old = atomicCAS(p, cur, __float_as_uint(sdata[0]));
            } while (old != cur);'''

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz,
                          [('NVARS', nvars), ('NEXPRS', nex),
                           ('FPMAX', f'FP({fpmax})')])}
{consts}

extern "C" __global__ void __launch_bounds__(256)
reduction({', '.join(args)})
{{
    __shared__ fpdtype_t sdata[256];
    fpdtype_t acc[NEXPRS];
    UNROLL for (int j = 0; j < NEXPRS; j++)
        acc[j] = {init};

    // One block of elements per CTA and turn: 32-bit index arithmetic
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x)
    {{
        const long long left = neles - blk*C_SUB;
        const int nvalid = (left < C_SUB) ? (int) left : C_SUB;

        for (int w = threadIdx.x; w < nrow*ld; w += blockDim.x)
        {{
            // storage word -> column -> (field variable, element)
            const long long i = blk*bsz + w;
            const int col = w % ld;
            const int v = (col / K_SOA) % NVARS;
            const int e = (col / (K_SOA*NVARS))*K_SOA + col % K_SOA;
            (void) v;

            if (e >= nvalid)
                continue;

            {acc}
        }}
    }}

    for (int j = 0; j < NEXPRS; j++)
    {{
        sdata[threadIdx.x] = acc[j];
        __syncthreads();

        for (int s = 128; s > 0; s >>= 1)
        {{
            if ((int) threadIdx.x < s)
                sdata[threadIdx.x] = {comb('sdata[threadIdx.x]',
                                           'sdata[threadIdx.x + s]')};
            __syncthreads();
        }}

        if (threadIdx.x == 0)
        {{
            {final}
        }}
        __syncthreads();
    }}
}}
'''
    return src, 'reduction', [x.split()[-1].lstrip('*') for x in args]
