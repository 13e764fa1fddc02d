"""Generator for the fused element kernel of the Navier-Stokes RHS.

One launch replaces the chain the reference runs as six to eight kernels
(``pyfr/solvers/baseadvecdiff/system.py:94-205``; the block-fusion group it
hands to ``Graph.group`` at ``:174-203``):

    tgradpcoru_upts   G  = (M4 - M6*M0) @ u
    tgradcoru_upts    G += M6 @ ucomm
    gradcoru_upts     G  = J^-T G                      (physical gradient)
    gradcoru_fpts     vect_fpts[d] = M0 @ G[d]
    tdisf             G  = S (Fi(u) + Fv(u, G))        (transformed flux)
    tdivtpcorf        fout = (M1 - M3*M2) @ G

``G`` (``ndims*nupts`` rows) never leaves the SM: per element block the
kernel reads ``u`` and the common solution, and writes only the gradients
at the flux points and the partial divergence -- ``(2 nupts + nfpts +
ndims nfpts)*LD`` words instead of the ``~35 nupts*LD`` the unfused chain
moves.

Per block (sm_100a):

* ``u`` and ``ucomm`` arrive by TMA bulk copy; the copy of the *next*
  block's ``ucomm`` is issued as soon as phase 1 has consumed the current
  one and the next ``u`` as soon as the flux has been formed, so HBM reads
  overlap the remaining phases;
* operator phases: a thread owns one column and a set of output rows; rows
  that read the same inputs (the points of one tensor-product line) are
  produced together from registers, which cuts shared-memory reads ~3x;
  operator constants are immediates;
* pointwise phases: a thread owns one (point, element) pair, exactly the
  arithmetic of the stand-alone ``gradcoru``/``tflux`` kernels.
"""

import numpy as np

from pyfr_b200.kernels import physics as ph
from pyfr_b200.kernels.mul import _pipeline_src


def _support(A, m):
    return tuple(np.flatnonzero(A[m]))


def _row_groups(terms, rows, smax=10):
    """Partition ``rows`` into groups reading identical inputs (across all
    terms); rows with a large or unique support stay on their own."""
    groups = {}

    for m in rows:
        key = tuple(_support(A, m) for A, _ in terms)
        nsup = sum(len(s) for s in key)
        if nsup > smax:
            key = ('solo', m)
        groups.setdefault(key, []).append(m)

    return list(groups.values())


def _balance(groups, R, cost):
    """Longest-processing-time assignment of groups to R bins."""
    bins, load = [[] for _ in range(R)], [0]*R

    for g in sorted(groups, key=cost, reverse=True):
        i = load.index(min(load))
        bins[i].append(g)
        load[i] += cost(g)

    return bins


def _fma_chain(pairs, acc=None):
    expr = acc
    for a, x in pairs:
        if expr is None:
            expr = f'{ph.fpconst(a)}*{x}'
        else:
            expr = f'fma({ph.fpconst(a)}, {x}, {expr})'
    return expr or 'FP(0.0)'


def emit_grouped(terms, R, store, LD):
    """Single-pass operator phase: every output row is finished inside the
    group that owns its inputs.  ``terms`` = [(A, smem array name)];
    ``store(m, expr)`` renders the store of output row ``m``."""
    M = terms[0][0].shape[0]
    groups = _row_groups(terms, range(M))
    cost = lambda g: sum(len(_support(A, g[0])) for A, _ in terms) + sum(
        int(np.count_nonzero(A[m])) for A, _ in terms for m in g)
    bins = _balance(groups, R, cost)

    cases = []
    for rg, gl in enumerate(bins):
        lines = []
        for g in gl:
            lines.append('{')
            regs = {}
            for ti, (A, src) in enumerate(terms):
                sup = sorted(set().union(*[_support(A, m) for m in g]))
                for k in sup:
                    regs[ti, k] = f'x{ti}_{k}'
                    lines.append(f'const fpdtype_t x{ti}_{k} = '
                                 f'{src}[{k*LD} + col];')
            for m in g:
                pairs = [(A[m, k], regs[ti, k])
                         for ti, (A, _) in enumerate(terms)
                         for k in _support(A, m)]
                lines.append(store(m, _fma_chain(pairs)))
            lines.append('}')

        cases.append(f'        case {rg}:\n            ' +
                     '\n            '.join(lines) + '\n            break;')

    return ('        switch (rg)\n        {\n' + '\n'.join(cases) +
            '\n        }\n')


def emit_accum(blocks, R, store, LD):
    """Operator phase whose rows gather from several column blocks
    (``blocks`` = [(A_d, smem name, row offset)]); a thread owns a
    contiguous run of rows, keeps their partial sums in registers and
    shares the loads of rows lying on one line of a block."""
    M = blocks[0][0].shape[0]
    bounds = np.linspace(0, M, R + 1).astype(int)

    cases = []
    for rg in range(R):
        rows = list(range(bounds[rg], bounds[rg + 1]))
        lines = [f'fpdtype_t a{j} = FP(0.0);' for j in range(len(rows))]

        for bi, (A, src, off) in enumerate(blocks):
            for g in _row_groups([(A, src)], rows):
                sup = sorted(set().union(*[_support(A, m) for m in g]))
                if not sup:
                    continue
                lines.append('{')
                for k in sup:
                    lines.append(f'const fpdtype_t x{k} = '
                                 f'{src}[{(off + k)*LD} + col];')
                for m in g:
                    j = rows.index(m)
                    pairs = [(A[m, k], f'x{k}') for k in _support(A, m)]
                    lines.append(f'a{j} = {_fma_chain(pairs, f"a{j}")};')
                lines.append('}')

        lines += [store(m, f'a{j}') for j, m in enumerate(rows)]
        cases.append(f'        case {rg}:\n            {{\n            ' +
                     '\n            '.join(lines) +
                     '\n            }\n            break;')

    return ('        switch (rg)\n        {\n' + '\n'.join(cases) +
            '\n        }\n')


def gradflux_source(be, ops, tplargs, pts, LD, R=None):
    """Source of the fused kernel.

    ``ops``: dict with the operator matrices ``A1`` (ndims*nupts x nupts),
    ``M6`` (ndims*nupts x nfpts), ``M0`` (nfpts x nupts) and ``A5``
    (nupts x ndims*nupts); ``tplargs``: the tflux template arguments
    (``ktype`` is 'linear' or 'curved')."""
    nd, nv = tplargs['ndims'], tplargs['nvars']
    A1, M6, M0, A5 = (np.asarray(ops[k], dtype=float)
                      for k in ('A1', 'M6', 'M0', 'A5'))
    nu, nf = M0.shape[1], M0.shape[0]
    isz = np.dtype(be.fpdtype).itemsize
    csub = be.csubsz

    assert A1.shape == (nd*nu, nu) and M6.shape == (nd*nu, nf)
    assert A5.shape == (nu, nd*nu) and LD == nv*csub

    wpr = -(-LD // 32)
    R = R or max(1, min(8, 16 // wpr))
    nthreads = 32*wpr*R
    linear = 'linear' in tplargs['ktype']

    smem = (nu + nf + nd*nu)*LD*isz + 64
    defs = [('NDIMS', nd), ('NVARS', nv), ('NPTS', nu), ('NFPTS', nf),
            ('NVERTS', tplargs.get('nverts', 0)), ('NEED_RCPDJAC', 1),
            ('LD', LD), ('NTHREADS', nthreads)]
    defs += ph.physics_defines(tplargs['c'], tplargs.get('visc_corr', 'none'),
                               True)

    # Phase 1: G = A1 @ U + M6 @ C
    p1 = emit_grouped([(A1, 'U'), (M6, 'C')], R,
                      lambda m, e: f'G[{m*LD} + col] = {e};', LD)

    # Phase 3: vect_fpts[d] = M0 @ G[d]
    M0d = np.zeros((nd*nf, nd*nu))
    for d in range(nd):
        M0d[d*nf:(d + 1)*nf, d*nu:(d + 1)*nu] = M0
    p3 = emit_grouped([(M0d, 'G')], R,
                      lambda m, e: f'vf[vfb + {m*LD} + col] = {e};', LD)

    # Phase 5: fout = A5 @ G, one column block per direction
    p5 = emit_accum([(A5[:, d*nu:(d + 1)*nu], 'G', d*nu) for d in range(nd)],
                    R, lambda m, e: f'fout[fob + {m*LD} + col] = {e};', LD)

    if linear:
        gsrc = ph.linear_smats_src(nd, tplargs['nverts'],
                                   tplargs['jac_exprs'])
        rows = ', '.join('{' + ', '.join(ph.fpconst(v) for v in row) + '}'
                         for row in pts)
        gsrc = (f'static __device__ const fpdtype_t c_pts[{len(pts)}][{nd}]'
                f' = {{{rows}}};\n' + gsrc)
        gargs = 'const fpdtype_t* __restrict__ verts, long long verts_bsz'
        geom = r'''
            fpdtype_t V[NVERTS][NDIMS], x[NDIMS], s[NDIMS][NDIMS], djac;
            UNROLL for (int n = 0; n < NVERTS; n++)
                UNROLL for (int i = 0; i < NDIMS; i++)
                    V[n][i] = __ldg(verts + blk*verts_bsz + n*(NDIMS*C_SUB)
                                    + COFF(e, i, NDIMS));
            UNROLL for (int i = 0; i < NDIMS; i++)
                x[i] = __ldg(&c_pts[p][i]);
            calc_smats_detj(V, x, s, djac);
            const fpdtype_t rcpdjac_v = FP(1.0)/djac;
'''
    else:
        gsrc = ''
        gargs = ('const fpdtype_t* __restrict__ smats, long long smats_bsz, '
                 'const fpdtype_t* __restrict__ rcpdjac, '
                 'long long rcpdjac_bsz')
        geom = r'''
            fpdtype_t s[NDIMS][NDIMS];
            UNROLL for (int i = 0; i < NDIMS; i++)
                UNROLL for (int j = 0; j < NDIMS; j++)
                    s[i][j] = __ldg(smats + blk*smats_bsz
                                    + (long long) (i*NPTS + p)*(NDIMS*C_SUB)
                                    + COFF(e, j, NDIMS));
            const fpdtype_t rcpdjac_v = __ldg(rcpdjac + blk*rcpdjac_bsz
                                              + p*C_SUB + e);
'''

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz, defs)}
{_pipeline_src}
{ph.flux_src}
{ph.visc_src}
{ph.geom_src}
{gsrc}

#define U_WORDS (NPTS*LD)
#define C_WORDS (NFPTS*LD)
#define G_WORDS (NDIMS*NPTS*LD)

extern "C" __global__ void __launch_bounds__(NTHREADS, 1)
gradflux(int nblocks, int neles,
         const fpdtype_t* __restrict__ u, long long u_bsz,
         const fpdtype_t* ucomm, long long ucomm_bsz,
         fpdtype_t* vf, long long vf_bsz,
         fpdtype_t* __restrict__ fout, long long fout_bsz,
         {gargs})
{{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    fpdtype_t *U = reinterpret_cast<fpdtype_t *>(smem_raw);
    fpdtype_t *C = U + U_WORDS;
    fpdtype_t *G = C + C_WORDS;
    unsigned long long *bars =
        reinterpret_cast<unsigned long long *>(G + G_WORDS);

    const int tid = threadIdx.x;
    const int col = tid % {32*wpr}, rg = tid / {32*wpr};
    const bool active = col < LD;

    if (tid == 0)
    {{
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }}
    __syncthreads();

    long long blk = blockIdx.x;

    if (tid == 0 && blk < nblocks)
    {{
        mbar_expect_tx(&bars[0], U_WORDS*sizeof(fpdtype_t));
        tma_load_1d(U, u + blk*u_bsz, U_WORDS*sizeof(fpdtype_t), &bars[0]);
        mbar_expect_tx(&bars[1], C_WORDS*sizeof(fpdtype_t));
        tma_load_1d(C, ucomm + blk*ucomm_bsz, C_WORDS*sizeof(fpdtype_t),
                    &bars[1]);
    }}

    for (unsigned it = 0; blk < nblocks; blk += gridDim.x, it++)
    {{
        const long long nxt = blk + gridDim.x;
        const long long vfb = blk*vf_bsz, fob = blk*fout_bsz;

        mbar_wait(&bars[0], it & 1);
        mbar_wait(&bars[1], it & 1);

        // ---- phase 1: corrected transformed gradient ------------------
        if (active)
        {{
{p1}
        }}
        __syncthreads();

        // ucomm consumed: fetch the next block's while we carry on
        if (tid == 0 && nxt < nblocks)
        {{
            mbar_expect_tx(&bars[1], C_WORDS*sizeof(fpdtype_t));
            tma_load_1d(C, ucomm + nxt*ucomm_bsz, C_WORDS*sizeof(fpdtype_t),
                        &bars[1]);
        }}

        // ---- phase 2: physical gradient (in place) ---------------------
        for (int item = tid; item < NPTS*C_SUB; item += NTHREADS)
        {{
            const int e = item % C_SUB, p = item / C_SUB;
            if (blk*C_SUB + e >= neles)
                continue;
{geom}
            fpdtype_t g[NDIMS][NVARS];
            UNROLL for (int d = 0; d < NDIMS; d++)
                UNROLL for (int v = 0; v < NVARS; v++)
                    g[d][v] = G[(d*NPTS + p)*LD + COFF(e, v, NVARS)];

            transform_grad(g, s, rcpdjac_v);

            UNROLL for (int d = 0; d < NDIMS; d++)
                UNROLL for (int v = 0; v < NVARS; v++)
                    G[(d*NPTS + p)*LD + COFF(e, v, NVARS)] = g[d][v];
        }}
        __syncthreads();

        // ---- phase 3: gradients at the flux points -> HBM ---------------
        if (active)
        {{
{p3}
        }}
        __syncthreads();

        // ---- phase 4: transformed flux (in place over the gradient) -----
        for (int item = tid; item < NPTS*C_SUB; item += NTHREADS)
        {{
            const int e = item % C_SUB, p = item / C_SUB;
            if (blk*C_SUB + e >= neles)
                continue;
{geom}
            (void) rcpdjac_v;
            fpdtype_t us[NVARS], g[NDIMS][NVARS];
            UNROLL for (int v = 0; v < NVARS; v++)
                us[v] = U[p*LD + COFF(e, v, NVARS)];
            UNROLL for (int d = 0; d < NDIMS; d++)
                UNROLL for (int v = 0; v < NVARS; v++)
                    g[d][v] = G[(d*NPTS + p)*LD + COFF(e, v, NVARS)];

            fpdtype_t ft[NDIMS][NVARS], fo[NDIMS][NVARS], pr, vel[NDIMS];
            inviscid_flux(us, ft, pr, vel);
            viscous_flux_add(us, g, ft);
            transform_flux(ft, s, fo);

            UNROLL for (int d = 0; d < NDIMS; d++)
                UNROLL for (int v = 0; v < NVARS; v++)
                    G[(d*NPTS + p)*LD + COFF(e, v, NVARS)] = fo[d][v];
        }}
        __syncthreads();

        // u consumed: fetch the next block's
        if (tid == 0 && nxt < nblocks)
        {{
            mbar_expect_tx(&bars[0], U_WORDS*sizeof(fpdtype_t));
            tma_load_1d(U, u + nxt*u_bsz, U_WORDS*sizeof(fpdtype_t),
                        &bars[0]);
        }}

        // ---- phase 5: divergence of the discontinuous flux -> HBM --------
        if (active)
        {{
{p5}
        }}
        __syncthreads();
    }}
}}
'''
    meta = dict(nthreads=nthreads, smem=smem, R=R,
                words_per_block=(2*nu + nf + nd*nf)*LD)

    return src, 'gradflux', meta
