"""Generator for the constant-operator multiply kernels.

Replaces the reference's GiMMiK / cuBLASLt ``mul`` providers
(``pyfr/backends/cuda/gimmik.py:23-121``, ``cublaslt.py:169-280``):
``out = alpha*A @ b + beta*out`` with ``A`` a constant ``M x K`` operator
(``M0``, ``M4 - M6*M0``, ``M6``, ``M1 - M3*M2``, ``M3``) and ``b``/``out``
blocked AoSoA matrices whose block ``j`` holds rows of ``LD`` columns
(``LD = nvars*csubsz``).

Design (sm_100a):

* ``A`` is baked into the instruction stream: only its non-zeros generate
  code and each becomes one DFMA/FFMA with an immediate-constant operand.
* persistent CTAs (one per SM) walk the element blocks; a block's input
  rows are contiguous in HBM, so the whole ``K x LD`` tile is fetched with
  a single TMA bulk copy (``cp.async.bulk``) into shared memory, double
  buffered and tracked with mbarriers so the fetch of block ``j+1``
  overlaps the arithmetic on block ``j``.
* a thread owns one column of the block and a contiguous group of output
  rows; shared-memory reads are conflict free (adjacent threads, adjacent
  words) and every output row is written as whole 128-byte segments.
* operators whose tile does not fit (``K = ndims*nupts``) are split into
  row chunks that stream through the same pipeline, with the partial
  sums held in registers.

Algorithmic HBM traffic per block is ``(K + M [+ M if beta != 0])*LD``
words: each input is read once, each output written once.
"""

import numpy as np

from pyfr_b200.kernels import physics as ph

_pipeline_src = r'''
__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return (unsigned) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned n)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(n));
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar,
                                               unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar,
                                          unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Barrier over the n threads (a multiple of 32) that name barrier id: lets
// the warp groups of a CTA synchronise independently of one another
__device__ __forceinline__ void bar_sync_named(int id, int n)
{
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory");
}

// One-shot flag in shared memory (release / acquire at CTA scope)
__device__ __forceinline__ void flag_set(int *f)
{
    asm volatile("st.release.cta.shared::cta.s32 [%0], 1;"
                 :: "r"(smem_u32(f)) : "memory");
}

__device__ __forceinline__ void flag_wait(int *f)
{
    int v;
    do
    {
        asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];"
                     : "=r"(v) : "r"(smem_u32(f)) : "memory");
    } while (!v);
}

// FP64 tensor-core product of an 8x4 by a 4x8 fragment, accumulated into
// this lane's two entries of the 8x8 result (lane = 4*g + t: a = A[g][t],
// b = B[t][g], c0/c1 = C[g][2t], C[g][2t + 1])
__device__ __forceinline__ void mma_m8n8k4(fpdtype_t &c0, fpdtype_t &c1,
                                           fpdtype_t a, fpdtype_t b)
{
#if PYFR_B200_FP64
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 "
                 "{%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
#endif
}

// One TMA bulk copy global -> shared, completion signalled on an mbarrier
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src,
                                            unsigned bytes,
                                            unsigned long long *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1], %2, [%3];"
        :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
'''

# (kept apart from ``_pipeline_src`` so that only the kernels that use it
# carry it)
_cpasync_src = r'''
// Per-thread asynchronous copy global -> shared of N = 4, 8 or 16 bytes
// (LDGSTS): gathers land in shared memory without passing through registers
template <int N>
__device__ __forceinline__ void cp_async(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;"
                 :: "r"(smem_u32(dst)), "l"(src), "n"(N) : "memory");
}

// ... 16 bytes, not kept in the first-level cache
__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                 :: "r"(smem_u32(dst)), "l"(src) : "memory");
}

// ... all of this thread's copies have landed (visible to the CTA after
// the next barrier)
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// ... or, without waiting: one arrival on an mbarrier once they have (the
// barrier's count includes these arrivals: .noinc)
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned long long *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}
'''


def plan_chunks(K, LD, itemsize, smem_budget, hint=None):
    """Split the K input rows into equal chunks whose double-buffered tile
    fits the shared-memory budget.  ``hint`` (e.g. nupts) is preferred as
    the chunk length when it divides K."""
    per_stage = smem_budget // 2
    maxrows = max(1, per_stage // (LD*itemsize))

    if K <= maxrows:
        return [(0, K)]

    cands = [d for d in range(1, K + 1) if K % d == 0 and d <= maxrows]
    rows = max(cands)
    if hint and K % hint == 0 and hint <= maxrows:
        rows = hint

    return [(i, i + rows) for i in range(0, K, rows)]


def mul_source(be, A, LD, alpha, beta, smem_budget=200*1024, rowgroups=4,
               chunk_hint=None, negdiv_nvars=None, rk=None):
    """CUDA source for ``out = alpha*A@b + beta*out``.

    With ``negdiv_nvars`` the ``negdivconf`` step that follows the last
    operator of the RHS (``out = -rcpdjac*out``, pyfr/solvers/baseadvec/
    kernels/negdivconf.mako) is applied in the epilogue, saving one full
    read-modify-write pass over the result.

    With ``rk`` (the ``rkvdh2`` template arguments, only together with
    ``negdiv_nvars``) the Runge-Kutta stage update that consumes the RHS
    (pyfr/integrators/explicit/kernels/rkvdh2.mako) is applied as well:
    the RHS value ``k`` never reaches memory, the kernel reads ``r1`` and
    writes ``r1 + dt a k`` and (into ``out``, the RHS bank ``r2``)
    ``r1 + dt b k``, plus the error-estimate registers when asked.

    Returns (source, name, launch meta dict)."""
    A = alpha*np.asarray(A, dtype=float)
    M, K = A.shape
    isz = np.dtype(be.fpdtype).itemsize

    if (LD*isz) % 16:
        raise ValueError('Row size must be a multiple of 16 bytes')

    # Coefficients: a 64-bit literal costs two uniform-register moves per
    # use unless it stays in a uniform register for the whole kernel, which
    # only works for operators with a handful of distinct values (hexes:
    # 6-13).  Dense operators (tets, pyramids: thousands of distinct
    # values; 1989 UMOV next to 2768 DFMA for the 90x56 one) can read them
    # as constant-bank operands of the FMA instead (opt-in until timed).
    uniq = np.unique(A[A != 0])
    ctab = (getattr(be, 'mul_const_table', 0) and isz == 8 and
            getattr(be, 'mul_const_table', 0) <= len(uniq) <= 7000)
    cidx = {float(v): i for i, v in enumerate(uniq)} if ctab else None
    coef = ((lambda a: f'KC[{cidx[float(a)]}]') if ctab else ph.fpconst)
    cdecl = (f'\n__constant__ fpdtype_t KC[{len(uniq)}] = {{'
             + ', '.join(ph.fpconst(v) for v in uniq) + '};') if ctab else ''

    chunks = plan_chunks(K, LD, isz, smem_budget, chunk_hint)
    nchunks = len(chunks)
    crows = chunks[0][1] - chunks[0][0]

    wpr = -(-LD // 32)                       # warps covering one row
    R = max(1, min(rowgroups, 32 // wpr, M))
    nthreads = 32*wpr*R

    # Narrow blocks (a small SoA width) leave most of an SM idle under one
    # CTA: run as many persistent CTAs per SM as fit 256 threads and the
    # shared memory
    smem_cta = 2*crows*LD*isz + 16
    nctas = max(1, min(256 // nthreads, smem_budget // (smem_cta + 1024)))

    # Output rows of each row group
    bounds = np.linspace(0, M, R + 1).astype(int)
    groups = [range(bounds[i], bounds[i + 1]) for i in range(R)]
    maxrows = max(len(g) for g in groups)

    def row_expr(m, k0, k1, acc=None):
        terms = [(k, A[m, k]) for k in range(k0, k1) if A[m, k] != 0]
        expr = acc

        for k, a in terms:
            t = f'sm[{(k - k0)*LD} + col]'
            if expr is None:
                expr = f'{coef(a)}*{t}'
            else:
                expr = f'fma({coef(a)}, {t}, {expr})'

        return expr

    if rk and not negdiv_nvars:
        raise ValueError('rk epilogue needs the negdivconf epilogue')

    def store(m, val):
        ix = f'ob + {m*LD} + col'
        if negdiv_nvars:
            old = f'out[{ix}] + ' if beta == 1 else (
                f'{ph.fpconst(beta)}*out[{ix}] + ' if beta else '')
            rhs = f'-__ldg(rcpdjac + rjb + {m}*C_SUB)*({old}{val})'

            if not rk:
                return f'out[{ix}] = {rhs};'

            # (everything this row reads from memory was fetched by
            # preload() at the top of the row group: with the loads written
            # next to the stores the compiler keeps them in program order
            # behind the previous row's stores, and the epilogue ran one
            # memory round trip per row -- 4.7 ms instead of 0.8, r02j)
            rhs = f'-rj_{m}*({"o_%d + " % m if beta else ""}{val})'
            st, last = rk['stage'], rk['stage'] == rk['nstages'] - 1
            c = lambda x: ph.fpconst(x[st])
            rix = lambda n: f'{n}[blk*{n}_bsz + {m*LD} + col]'
            L = [f'{{ const fpdtype_t kk = {rhs}, t1 = t_{m};']
            if rk['errest'] and st == 0:
                L += [f'{rix("rerr")} = dt*{c(rk["e"])}*kk;',
                      f'{rix("rold")} = t1;']
            elif rk['errest']:
                L += [f'{rix("rerr")} = e_{m} + dt*{c(rk["e"])}*kk;']
            if last:
                L += [f'{rix("r1")} = t1 + dt*{c(rk["b"])}*kk; }}']
            else:
                L += [f'{rix("r1")} = t1 + dt*{c(rk["a"])}*kk;',
                      f'out[{ix}] = t1 + dt*{c(rk["b"])}*kk; }}']
            return ' '.join(L)
        if beta == 0:
            return f'out[{ix}] = {val};'
        elif beta == 1:
            return f'out[{ix}] += {val};'
        else:
            return f'out[{ix}] = fma({ph.fpconst(beta)}, out[{ix}], {val});'

    def preload(m):
        """Loads of the fused stage update of row ``m``, hoisted."""
        ix = f'ob + {m*LD} + col'
        rix = lambda n: f'{n}[blk*{n}_bsz + {m*LD} + col]'
        L = [f'rj_{m} = __ldg(rcpdjac + rjb + {m}*C_SUB)',
             f't_{m} = {rix("r1")}']
        if beta:
            L.append(f'o_{m} = ' + ('' if beta == 1 else
                                    f'{ph.fpconst(beta)}*') + f'out[{ix}]')
        if rk['errest'] and rk['stage'] > 0:
            L.append(f'e_{m} = {rix("rerr")}')
        return 'const fpdtype_t ' + ', '.join(L) + ';'

    cases = []
    for ci, (k0, k1) in enumerate(chunks):
        first, last = ci == 0, ci == nchunks - 1
        body = []

        for rg, rows in enumerate(groups):
            lines = []
            if rk and last:
                lines += [preload(m) for m in rows]
            for j, m in enumerate(rows):
                if nchunks == 1:
                    e = row_expr(m, k0, k1) or 'FP(0.0)'
                    lines.append(store(m, e))
                else:
                    e = row_expr(m, k0, k1, None if first else f'acc[{j}]')
                    if e is None:
                        e = 'FP(0.0)'
                    if last:
                        lines.append(store(m, e))
                    elif e != f'acc[{j}]':
                        lines.append(f'acc[{j}] = {e};')

            body.append(f'            case {rg}:\n            {{\n'
                        '                ' +
                        '\n                '.join(lines) +
                        '\n                break;\n            }')

        cases.append(f'        case {ci}:\n            switch (rg)\n'
                     '            {\n' + '\n'.join(body) +
                     '\n            }\n            break;')

    extra_args = extra_pre = ''
    if negdiv_nvars:
        extra_args = (', const fpdtype_t* __restrict__ rcpdjac, '
                      'long long rcpdjac_bsz')
        extra_pre = (f'        const long long rjb = blk*rcpdjac_bsz + '
                     f'(col/(K_SOA*{negdiv_nvars}))*K_SOA + col % K_SOA;')
    if rk:
        regs = ['r1'] + (['rold', 'rerr'] if rk['errest'] else [])
        extra_args += ''.join(f', fpdtype_t* __restrict__ {n}, '
                              f'long long {n}_bsz' for n in regs)
        extra_args += ', const fpdtype_t* __restrict__ dt_p'

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz)}
#define LD {LD}
#define NCHUNKS {nchunks}
#define CROWS {crows}
#define TILE (CROWS*LD)
#define NTHREADS {nthreads}
{_pipeline_src}{cdecl}

// out[{M} x LD] = A[{M} x {K}] @ b[{K} x LD] per element block;
// {int(np.count_nonzero(A))} non-zeros, {nchunks} chunk(s) of {crows} rows
extern "C" __global__ void __launch_bounds__(NTHREADS, {nctas})
opmul(int nblocks, const fpdtype_t* __restrict__ b, long long b_bsz,
      fpdtype_t* __restrict__ out, long long out_bsz{extra_args})
{{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    fpdtype_t *tiles = reinterpret_cast<fpdtype_t *>(smem_raw);
    unsigned long long *full =
        reinterpret_cast<unsigned long long *>(tiles + 2*TILE);

    const int tid = threadIdx.x;
    const int col = tid % {32*wpr}, rg = tid / {32*wpr};
    {'const fpdtype_t dt = *dt_p;' if rk else ''}
    const bool active = col < LD;

    // Work items: (block, chunk) pairs owned by this CTA, in order
    const long long myblocks = (nblocks - (long long) blockIdx.x
                                + gridDim.x - 1) / gridDim.x;
    const long long nitems = myblocks*NCHUNKS;

    if (tid == 0)
    {{
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }}
    __syncthreads();

    auto issue = [&](long long item)
    {{
        const long long blk = blockIdx.x + (item / NCHUNKS)*gridDim.x;
        const int chunk = (int) (item % NCHUNKS), st = (int) (item & 1);
        const fpdtype_t *src = b + blk*b_bsz + (long long) chunk*TILE;

        mbar_expect_tx(&full[st], TILE*sizeof(fpdtype_t));
        tma_load_1d(tiles + st*TILE, src, TILE*sizeof(fpdtype_t), &full[st]);
    }};

    if (tid == 0 && nitems > 0)
        issue(0);

    fpdtype_t acc[{maxrows if nchunks > 1 else 1}];

    for (long long item = 0; item < nitems; item++)
    {{
        const int st = (int) (item & 1), chunk = (int) (item % NCHUNKS);
        const long long blk = blockIdx.x + (item / NCHUNKS)*gridDim.x;

        // Prefetch the next tile into the buffer released last iteration
        if (tid == 0 && item + 1 < nitems)
            issue(item + 1);

        mbar_wait(&full[st], (unsigned) ((item >> 1) & 1));

        const fpdtype_t *sm = tiles + st*TILE;
        const long long ob = blk*out_bsz;
{extra_pre}

        if (active)
        {{
        switch (chunk)
        {{
{chr(10).join(cases)}
        }}
        }}

        __syncthreads();
    }}
}}
'''

    meta = dict(nthreads=nthreads, nctas=nctas,
                smem=2*crows*LD*isz + 16, nnz=int(np.count_nonzero(A)),
                nchunks=nchunks, crows=crows, M=M, K=K)

    return src, 'opmul', meta
