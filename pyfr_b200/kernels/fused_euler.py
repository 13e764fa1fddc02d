"""Generator for the fused element kernel of the Euler RHS.

One launch replaces the second graph of the reference's advection systems
(``pyfr/solvers/baseadvec/system.py:93-136`` and the block-fusion group it
hands to ``Graph.group`` at ``:126-134``):

    tdisf        F  = S Fi(u)                       (transformed flux)
    tdivtpcorf   r  = (M1 - M3*M2) @ F
    tdivtconf    r += M3 @ fcomm                    (common-flux correction)
    negdivconf   r  = -r / |J|

``F`` (``ndims*nupts`` rows per block) never leaves the SM: the kernel reads
``u`` and the common normal fluxes at the flux points and writes the
negated divergence -- ``(2 nupts + nfpts)*LD`` words per block instead of
the ``(4 + 2 ndims) nupts + nfpts`` of the three separate launches.

Same building blocks as the Navier-Stokes kernel (``fused.py``): TMA bulk
copies tracked by mbarriers (``u`` and the vertices are refetched as soon
as the flux phase has consumed them, the common fluxes as soon as the
correction phase has), operators applied as table-driven line classes with
baked coefficients, pointwise phase identical to the stand-alone ``tflux``.
"""

import numpy as np

from pyfr_b200.kernels import physics as ph
from pyfr_b200.kernels.fused import (ConstPool, NotFusable, PhaseEmitter,
                                     build_classes)
from pyfr_b200.kernels.mul import _pipeline_src


def fluxdiv_source(be, ops, tplargs, pts, LD, rk=None):
    """``ops``: ``A5`` (nupts x ndims*nupts) and ``M3`` (nupts x nfpts);
    ``tplargs``: the Euler ``tflux`` template arguments (``ktype`` 'linear'
    or 'curved').  ``rk``: template arguments of an ``rkvdh2`` stage to
    apply to the result on its way out (see ``mul_source``).  Raises
    ``NotFusable`` when the operators lack line structure or the block does
    not fit shared memory."""
    nd, nv = tplargs['ndims'], tplargs['nvars']
    A5, M3 = (np.asarray(ops[k], dtype=float) for k in ('A5', 'M3'))
    nu, nf = M3.shape
    isz = np.dtype(be.fpdtype).itemsize
    csub = be.csubsz

    if A5.shape != (nu, nd*nu) or LD != nv*csub or be.soasz != csub:
        raise NotFusable('unexpected operator shapes / layout')

    linear = 'linear' in tplargs['ktype']
    npoints = nu*csub
    nthreads = 512 if npoints >= 512 else max(128, -(-npoints // 32)*32)
    nverts = tplargs.get('nverts', 0)

    defs = [('NDIMS', nd), ('NVARS', nv), ('NPTS', nu), ('NFPTS', nf),
            ('NVERTS', nverts), ('LD', LD), ('NTHREADS', nthreads)]
    defs += ph.physics_defines(tplargs['c'])

    K = ConstPool(isz == 8)
    em = PhaseEmitter(LD, isz, K)

    # Divergence: in-place line transforms, direction by direction
    A5d = np.zeros((nd*nu, nd*nu))
    for d in range(nd):
        A5d[d*nu:(d + 1)*nu, d*nu:(d + 1)*nu] = A5[:, d*nu:(d + 1)*nu]
    p5 = em.emit('p5', build_classes([A5d]), ['G'],
                 lambda off, v, at: f'{at("G", off)} = {v};', inplace=True)

    # Correction: rows 0..nupts of G accumulate M3 @ fcomm
    pm = em.emit('pm', build_classes([M3]), ['C'],
                 lambda off, v, at: f'{at("G", off)} += {v};')

    psum = ' + '.join(f'G[{d*nu*LD} + item]' for d in range(nd))

    if linear:
        rows = ', '.join(ph.fpconst(v) for row in pts for v in row)
        gsrc = (f'static __device__ const fpdtype_t c_pts[{len(pts)*nd}] = '
                f'{{{rows}}};\n'
                + ph.linear_smats_src(nd, nverts, tplargs['jac_exprs']))
        gargs = 'const fpdtype_t* __restrict__ verts, long long verts_bsz'
        geom = r'''
                fpdtype_t V[NVERTS][NDIMS], x[NDIMS], s[NDIMS][NDIMS], djac;
                UNROLL for (int n = 0; n < NVERTS; n++)
                    UNROLL for (int i = 0; i < NDIMS; i++)
                        V[n][i] = VS[n*(NDIMS*C_SUB) + COFF(e, i, NDIMS)];
                UNROLL for (int i = 0; i < NDIMS; i++)
                    x[i] = PTS[p*NDIMS + i];
                calc_smats_detj(V, x, s, djac);
                const fpdtype_t rcpdjac_v = FP(1.0)/djac;
'''
        v_words, pt_words = nverts*nd*csub, -(-nu*nd // 4)*4
        geo_decl = f'''fpdtype_t *VS = RJ + {npoints};
    fpdtype_t *PTS = VS + {v_words};'''
        geo_stage = f'''for (int i = tid; i < {nu*nd}; i += NTHREADS)
        PTS[i] = c_pts[i];'''
        geo_fetch = ('tma_load_1d(VS, verts + b*verts_bsz, '
                     f'{v_words}*sizeof(fpdtype_t), &bars[0]);')
        geo_bytes = f' + {v_words}*sizeof(fpdtype_t)'
        geo_words = v_words + pt_words
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
        geo_decl = geo_stage = geo_fetch = geo_bytes = ''
        geo_words = 0

    tables = em.decls()
    fixed = ((nu + nf + nd*nu)*LD + npoints + geo_words)*isz + 64
    smem_sm = 228*1024
    nctas = max(1, min(512 // nthreads, smem_sm // (fixed + 1024)))
    em.plan(min(smem_sm // nctas - 1024, 227*1024) - fixed)
    smem = fixed + em.table_bytes

    if smem > 227*1024:
        raise NotFusable(f'needs {smem} bytes of shared memory')

    if rk:
        st, last = rk['stage'], rk['stage'] == rk['nstages'] - 1
        c = lambda x: ph.fpconst(x[st])
        rix = lambda n: f'{n}[blk*{n}_bsz + item]'
        L = [f'const fpdtype_t kk = -RJ[p*C_SUB + e]*({psum}), '
             f't1 = {rix("r1")};']
        if rk['errest'] and st == 0:
            L += [f'{rix("rerr")} = dt*{c(rk["e"])}*kk;',
                  f'{rix("rold")} = t1;']
        elif rk['errest']:
            L += [f'{rix("rerr")} = {rix("rerr")} + dt*{c(rk["e"])}*kk;']
        if last:
            L += [f'{rix("r1")} = t1 + dt*{c(rk["b"])}*kk;']
        else:
            L += [f'{rix("r1")} = t1 + dt*{c(rk["a"])}*kk;',
                  f'fout[fob + item] = t1 + dt*{c(rk["b"])}*kk;']
        out_stmt = '\n            '.join(L)
        regs = ['r1'] + (['rold', 'rerr'] if rk['errest'] else [])
        rkargs = ''.join(f',\n        fpdtype_t* __restrict__ {n}, '
                         f'long long {n}_bsz' for n in regs)
        rkargs += ',\n        const fpdtype_t* __restrict__ dt_p'
    else:
        out_stmt = f'fout[fob + item] = -RJ[p*C_SUB + e]*({psum});'
        rkargs = ''

    tail = f'RJ + {npoints} + {geo_words}'

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz, defs)}
{_pipeline_src}
{ph.flux_src}
{ph.geom_src}
{gsrc}
{tables}
{K.decl()}

#define U_WORDS (NPTS*LD)
#define C_WORDS (NFPTS*LD)
#define G_WORDS (NDIMS*NPTS*LD)

extern "C" __global__ void __launch_bounds__(NTHREADS, {nctas})
fluxdiv(int nblocks, int neles,
        const fpdtype_t* __restrict__ u, long long u_bsz,
        const fpdtype_t* __restrict__ fcomm, long long fcomm_bsz,
        fpdtype_t* __restrict__ fout, long long fout_bsz,
        {gargs}{rkargs})
{{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    fpdtype_t *U = reinterpret_cast<fpdtype_t *>(smem_raw);
    fpdtype_t *C = U + U_WORDS;
    fpdtype_t *G = C + C_WORDS;
    fpdtype_t *RJ = G + G_WORDS;
    {geo_decl}
    {em.smem_layout(tail)}
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(
        reinterpret_cast<unsigned char *>({tail}) + {em.table_bytes});

    const int tid = threadIdx.x;
    {'const fpdtype_t dt = *dt_p;' if rk else ''}

    {em.stage(tail)}
    {geo_stage}

    if (tid == 0)
    {{
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }}
    __syncthreads();

    auto fetch_u = [&](long long b)
    {{
        mbar_expect_tx(&bars[0], U_WORDS*sizeof(fpdtype_t){geo_bytes});
        tma_load_1d(U, u + b*u_bsz, U_WORDS*sizeof(fpdtype_t), &bars[0]);
        {geo_fetch}
    }};
    auto fetch_c = [&](long long b)
    {{
        mbar_expect_tx(&bars[1], C_WORDS*sizeof(fpdtype_t));
        tma_load_1d(C, fcomm + b*fcomm_bsz, C_WORDS*sizeof(fpdtype_t),
                    &bars[1]);
    }};

    long long blk = blockIdx.x;
    if (tid == 0 && blk < nblocks)
    {{
        fetch_u(blk);
        fetch_c(blk);
    }}

    for (unsigned it = 0; blk < nblocks; blk += gridDim.x, it++)
    {{
        const long long nxt = blk + gridDim.x;
        const long long fob = blk*fout_bsz;

        // ---- transformed flux at the solution points ---------------------
        mbar_wait(&bars[0], it & 1);

        for (int item = tid; item < NPTS*C_SUB; item += NTHREADS)
        {{
            const int e = item % C_SUB, p = item / C_SUB;

            if (blk*C_SUB + e < neles)
            {{
{geom}
                fpdtype_t us[NVARS], ft[NDIMS][NVARS], fo[NDIMS][NVARS];
                fpdtype_t pr, vel[NDIMS];
                UNROLL for (int v = 0; v < NVARS; v++)
                    us[v] = U[p*LD + COFF(e, v, NVARS)];

                inviscid_flux(us, ft, pr, vel);
                transform_flux(ft, s, fo);

                UNROLL for (int d = 0; d < NDIMS; d++)
                    UNROLL for (int v = 0; v < NVARS; v++)
                        G[(d*NPTS + p)*LD + COFF(e, v, NVARS)] = fo[d][v];
                RJ[item] = rcpdjac_v;
            }}
            else
            {{
                UNROLL for (int d = 0; d < NDIMS; d++)
                    UNROLL for (int v = 0; v < NVARS; v++)
                        G[(d*NPTS + p)*LD + COFF(e, v, NVARS)] = FP(0.0);
                RJ[item] = FP(0.0);
            }}
        }}
        __syncthreads();

        // u (and the vertices) are consumed: fetch the next block's
        if (tid == 0 && nxt < nblocks)
            fetch_u(nxt);

        // ---- divergence: line transforms per direction --------------------
        {p5}
        __syncthreads();

        // ---- common-flux correction into the first direction's rows -------
        mbar_wait(&bars[1], it & 1);
        {pm}
        __syncthreads();

        if (tid == 0 && nxt < nblocks)
            fetch_c(nxt);

        // ---- sum of the directions, physical scaling, out -----------------
        for (int item = tid; item < NPTS*LD; item += NTHREADS)
        {{
            const int p = item / LD, col = item - p*LD;
            const int e = (col / (K_SOA*NVARS))*K_SOA + col % K_SOA;
            {out_stmt}
        }}
        __syncthreads();
    }}
}}
'''
    meta = dict(nthreads=nthreads, smem=smem, nctas=nctas,
                words_per_block=(2*nu + nf)*LD)

    return src, 'fluxdiv', meta
