"""CUDA source for the element-local pointwise kernels.

``tflux`` (Euler: ``pyfr/solvers/euler/kernels/tflux.mako``; Navier-Stokes:
``pyfr/solvers/navstokes/kernels/tflux.mako``), ``gradcoru``
(``pyfr/solvers/baseadvecdiff/kernels/gradcoru.mako``) and ``negdivconf``
(``pyfr/solvers/baseadvec/kernels/negdivconf.mako``).

Work decomposition (blocked AoSoA, one block = ``C_SUB`` elements): a
thread owns one (solution point, element) pair; consecutive threads walk
the ``C_SUB`` elements of a row and then move to the next point, so every
load/store instruction of a warp touches whole 128-byte row segments.  All
``NVARS`` (and ``NDIMS x NVARS``) values of the point live in registers.
Grids are sized in whole blocks; one launch covers every element of the
region.
"""

import re

from pyfr_b200.kernels import physics as ph

# Thread <-> (block, point, element-in-block) decomposition shared by the
# 2-D kernels; returns early for the padding columns of the last block
_index_src = r'''
    const long long gid = (long long) blockIdx.x*blockDim.x + threadIdx.x;
    const int e = (int) (gid % C_SUB);
    const int p = (int) ((gid / C_SUB) % NPTS);
    const long long blk = gid / ((long long) C_SUB*NPTS);

    if (blk*C_SUB + e >= neles)
        return;
'''

_geom_linear_src = r'''
    // Metric terms of a linear element, rebuilt from its vertices
    fpdtype_t V[NVERTS][NDIMS], x[NDIMS], s[NDIMS][NDIMS], djac;
    UNROLL for (int n = 0; n < NVERTS; n++)
        UNROLL for (int i = 0; i < NDIMS; i++)
            V[n][i] = __ldg(verts + blk*verts_bsz + n*(NDIMS*C_SUB)
                            + COFF(e, i, NDIMS));
    UNROLL for (int i = 0; i < NDIMS; i++)
        x[i] = __ldg(&c_pts[p][i]);

    calc_smats_detj(V, x, s, djac);
    const fpdtype_t rcpdjac_v = FP(1.0)/djac;
    (void) rcpdjac_v;
'''

_geom_curved_src = r'''
    // Stored metric terms of a curved element
    fpdtype_t s[NDIMS][NDIMS];
    UNROLL for (int i = 0; i < NDIMS; i++)
        UNROLL for (int j = 0; j < NDIMS; j++)
            s[i][j] = __ldg(smats + blk*smats_bsz
                            + (long long) (i*NPTS + p)*(NDIMS*C_SUB)
                            + COFF(e, j, NDIMS));
#ifdef NEED_RCPDJAC
    const fpdtype_t rcpdjac_v = __ldg(rcpdjac + blk*rcpdjac_bsz + p*C_SUB + e);
#endif
'''


def _pts_table(pts):
    rows = ', '.join('{' + ', '.join(ph.fpconst(v) for v in row) + '}'
                     for row in pts)
    return (f'static __device__ const fpdtype_t c_pts[{len(pts)}]'
            f'[{len(pts[0])}] = {{{rows}}};\n')


def _geom(tplargs, pts):
    if 'linear' in tplargs['ktype']:
        src = (_pts_table(pts) +
               ph.linear_smats_src(tplargs['ndims'], tplargs['nverts'],
                                   tplargs['jac_exprs']))
        args = ['const fpdtype_t* __restrict__ verts', 'long long verts_bsz']
        return src, args, _geom_linear_src
    else:
        args = ['const fpdtype_t* __restrict__ smats', 'long long smats_bsz',
                'const fpdtype_t* __restrict__ rcpdjac',
                'long long rcpdjac_bsz']
        return '', args, _geom_curved_src


def tflux_source(be, tplargs, npts, pts, viscous):
    """Returns (source, kernel name, ordered argument names)."""
    nd, nv = tplargs['ndims'], tplargs['nvars']
    fused = 'fused' in tplargs['ktype']

    defs = [('NDIMS', nd), ('NVARS', nv), ('NPTS', npts),
            ('NVERTS', tplargs.get('nverts', 0))]
    defs += ph.physics_defines(tplargs['c'], tplargs.get('visc_corr', 'none'),
                               viscous)
    if fused:
        defs.append(('NEED_RCPDJAC', 1))

    gsrc, gargs, gbody = _geom(tplargs, pts)

    args = ['int neles', 'const fpdtype_t* __restrict__ u', 'long long u_bsz',
            'fpdtype_t* __restrict__ f', 'long long f_bsz']
    if fused:
        args += ['fpdtype_t* __restrict__ gradu', 'long long gradu_bsz']
    args += gargs

    if viscous and fused:
        load_grad = '''
    // Corrected transformed gradient -> physical gradient, written back
    fpdtype_t g[NDIMS][NVARS];
    UNROLL for (int d = 0; d < NDIMS; d++)
        UNROLL for (int v = 0; v < NVARS; v++)
            g[d][v] = gradu[blk*gradu_bsz + (long long) (d*NPTS + p)*LD + COFF(e, v, NVARS)];

    transform_grad(g, s, rcpdjac_v);

    UNROLL for (int d = 0; d < NDIMS; d++)
        UNROLL for (int v = 0; v < NVARS; v++)
            gradu[blk*gradu_bsz + (long long) (d*NPTS + p)*LD + COFF(e, v, NVARS)] = g[d][v];
'''
    elif viscous:
        load_grad = '''
    // The flux buffer holds the physical gradient on entry
    fpdtype_t g[NDIMS][NVARS];
    UNROLL for (int d = 0; d < NDIMS; d++)
        UNROLL for (int v = 0; v < NVARS; v++)
            g[d][v] = f[blk*f_bsz + (long long) (d*NPTS + p)*LD + COFF(e, v, NVARS)];
'''
    else:
        load_grad = ''

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz, defs)}
#define LD (NVARS*C_SUB)
{ph.flux_src}
{ph.visc_src if viscous else ''}
{ph.geom_src}
{gsrc}

extern "C" __global__ void __launch_bounds__(128)
tflux({', '.join(args)})
{{
{_index_src}
{gbody}

    fpdtype_t us[NVARS];
    UNROLL for (int v = 0; v < NVARS; v++)
        us[v] = u[blk*u_bsz + (long long) p*LD + COFF(e, v, NVARS)];
{load_grad}
    // Physical flux F = Fi (+ Fv), then its contravariant transform
    fpdtype_t ft[NDIMS][NVARS], fo[NDIMS][NVARS], pr, vel[NDIMS];
    inviscid_flux(us, ft, pr, vel);
    {'viscous_flux_add(us, g, ft);' if viscous else ''}
    transform_flux(ft, s, fo);

    UNROLL for (int d = 0; d < NDIMS; d++)
        UNROLL for (int v = 0; v < NVARS; v++)
            f[blk*f_bsz + (long long) (d*NPTS + p)*LD + COFF(e, v, NVARS)] = fo[d][v];
}}
'''
    return src, 'tflux', [a.split()[-1].lstrip('*') for a in args]


def gradcoru_source(be, tplargs, npts, pts):
    nd, nv = tplargs['ndims'], tplargs['nvars']
    defs = [('NDIMS', nd), ('NVARS', nv), ('NPTS', npts),
            ('NVERTS', tplargs.get('nverts', 0)), ('NEED_RCPDJAC', 1)]

    gsrc, gargs, gbody = _geom(tplargs, pts)
    args = ['int neles', 'fpdtype_t* __restrict__ gradu',
            'long long gradu_bsz'] + gargs

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz, defs)}
#define LD (NVARS*C_SUB)
{ph.geom_src}
{gsrc}

extern "C" __global__ void __launch_bounds__(128)
gradcoru({', '.join(args)})
{{
{_index_src}
{gbody}

    fpdtype_t g[NDIMS][NVARS];
    UNROLL for (int d = 0; d < NDIMS; d++)
        UNROLL for (int v = 0; v < NVARS; v++)
            g[d][v] = gradu[blk*gradu_bsz + (long long) (d*NPTS + p)*LD + COFF(e, v, NVARS)];

    transform_grad(g, s, rcpdjac_v);

    UNROLL for (int d = 0; d < NDIMS; d++)
        UNROLL for (int v = 0; v < NVARS; v++)
            gradu[blk*gradu_bsz + (long long) (d*NPTS + p)*LD + COFF(e, v, NVARS)] = g[d][v];
}}
'''
    return src, 'gradcoru', [a.split()[-1].lstrip('*') for a in args]


def negdivconf_source(be, tplargs, npts):
    defs = [('NDIMS', tplargs['ndims']), ('NVARS', tplargs['nvars']),
            ('NPTS', npts)]
    args = ['int neles', 'fpdtype_t* __restrict__ tdivtconf',
            'long long tdivtconf_bsz', 'const fpdtype_t* __restrict__ rcpdjac',
            'long long rcpdjac_bsz']

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz, defs)}
#define LD (NVARS*C_SUB)

extern "C" __global__ void __launch_bounds__(128)
negdivconf({', '.join(args)})
{{
{_index_src}
    const fpdtype_t r = -__ldg(rcpdjac + blk*rcpdjac_bsz + p*C_SUB + e);

    UNROLL for (int v = 0; v < NVARS; v++)
    {{
        const long long ix = blk*tdivtconf_bsz + (long long) p*LD + COFF(e, v, NVARS);
        tdivtconf[ix] = r*tdivtconf[ix];
    }}
}}
'''
    return src, 'negdivconf', [a.split()[-1].lstrip('*') for a in args]


def fieldeval_source(be, tplargs, npts):
    """``fieldeval`` (``pyfr/plugins/kernels/fieldeval.mako``): per
    element, the weighted sum -- or the minimum / maximum, optionally under
    a mask -- over its points of each expression in the primitive
    variables (``pri[i]``), their physical gradients (``grad_pri[i][d]``,
    ``con_to_pri``/``grad_con_to_pri`` of
    ``pyfr/solvers/euler/kernels/eos.mako:17-45``), the coordinates
    (``ploc[d]``) and the time ``t``.

    The expressions arrive as C text (the reference compiles them the same
    way, ``pyfr/plugins/fieldeval.py:12-22``) and are pasted into the
    kernel.  One thread owns one element and walks its points, so the
    per-element result needs no inter-thread reduction."""
    nd, nv = tplargs['ndims'], tplargs['nvars']
    exprs, grads = tplargs['exprs'], tplargs['has_grads']

    # The integrate plugin wraps L-p integrands in abs(): make sure the
    # floating-point function is the one that is called
    exprs = [re.sub(r'\babs\(', 'fabs(', e) for e in exprs]

    rop, has_wts = tplargs['reduceop'], tplargs.get('has_wts', True)
    has_ploc = tplargs.get('has_ploc', False)

    if tplargs.get('use_views') or rop not in ('sum', 'min', 'max'):
        raise NotImplementedError('fieldeval over interface views is not on '
                                  'the b200 path')
    if rop == 'sum' and not has_wts:
        raise ValueError('fieldeval: a sum needs weights')

    fpmax = ('1.7976931348623157e308' if be.fpdtype.__name__ == 'float64'
             else '3.4028234e38f')
    init = {'sum': 'FP(0.0)', 'max': f'-FP({fpmax})',
            'min': f'FP({fpmax})'}[rop]

    defs = [('NDIMS', nd), ('NVARS', nv), ('NPTS', npts),
            ('NEXPRS', len(exprs)),
            ('C_GM1', ph.fpconst(tplargs['c']['gamma'] - 1))]

    args = ['int neles', 'const fpdtype_t* __restrict__ u', 'long long u_bsz']
    if grads:
        args += ['const fpdtype_t* __restrict__ gradu', 'long long gradu_bsz']
    if has_ploc:
        args += ['const fpdtype_t* __restrict__ plocm', 'long long ploc_bsz']
    if has_wts:
        args += ['const fpdtype_t* __restrict__ wts', 'long long wts_bsz',
                 'int wts_ld']
    args += ['fpdtype_t* __restrict__ out', 'long long out_bsz', 'int out_ld',
             'const fpdtype_t* __restrict__ t_p']

    gsrc = r'''
        fpdtype_t grad_pri[NVARS][NDIMS];
        {
            fpdtype_t gc[NDIMS][NVARS], vel[NDIMS], rhov[NDIMS];
            UNROLL for (int d = 0; d < NDIMS; d++)
                UNROLL for (int v = 0; v < NVARS; v++)
                    gc[d][v] = gradu[blk*gradu_bsz
                                     + (long long) (d*NPTS + p)*LD
                                     + COFF(e, v, NVARS)];
            UNROLL for (int i = 0; i < NDIMS; i++)
            {
                rhov[i] = cons[i + 1];
                vel[i] = invrho*rhov[i];
            }
            UNROLL for (int d = 0; d < NDIMS; d++)
                grad_pri[0][d] = gc[d][0];
            UNROLL for (int i = 0; i < NDIMS; i++)
                UNROLL for (int d = 0; d < NDIMS; d++)
                    grad_pri[i + 1][d] = invrho*(gc[d][i + 1]
                                                 - vel[i]*gc[d][0]);
            UNROLL for (int d = 0; d < NDIMS; d++)
            {
                fpdtype_t term = 0;
                UNROLL for (int i = 0; i < NDIMS; i++)
                    term += vel[i]*gc[d][i + 1] + rhov[i]*grad_pri[i + 1][d];
                grad_pri[NVARS - 1][d] = C_GM1*(gc[d][NVARS - 1]
                                                - FP(0.5)*term);
            }
        }
''' if grads else ''

    if rop == 'sum':
        esrc = '\n'.join(f'        acc[{j}] += w*({e});'
                         for j, e in enumerate(exprs))
    else:
        # the weights of a min / max are a mask: points with w <= 0 drop out
        fn = 'fmax' if rop == 'max' else 'fmin'
        esrc = '\n'.join(
            f'        acc[{j}] = {fn}(acc[{j}], '
            + (f'(w > 0) ? ({e}) : {init}' if has_wts else f'({e})') + ');'
            for j, e in enumerate(exprs)
        )

    psrc = '''
        fpdtype_t ploc[NDIMS];
        UNROLL for (int d = 0; d < NDIMS; d++)
            ploc[d] = __ldg(plocm + blk*ploc_bsz
                            + (long long) p*(NDIMS*C_SUB) + COFF(e, d, NDIMS));
''' if has_ploc else ''

    wsrc = ('const fpdtype_t w = __ldg(wts + blk*wts_bsz + '
            '(long long) p*wts_ld + e);' if has_wts else '')

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz, defs)}
#define LD (NVARS*C_SUB)

extern "C" __global__ void __launch_bounds__(128)
fieldeval({', '.join(args)})
{{
    const long long gid = (long long) blockIdx.x*blockDim.x + threadIdx.x;
    const int e = (int) (gid % C_SUB);
    const long long blk = gid / C_SUB;

    if (blk*C_SUB + e >= neles)
        return;

    const fpdtype_t t = *t_p;
    fpdtype_t acc[NEXPRS];
    UNROLL for (int j = 0; j < NEXPRS; j++)
        acc[j] = {init};

    for (int p = 0; p < NPTS; p++)
    {{
        fpdtype_t cons[NVARS], pri[NVARS];
        UNROLL for (int v = 0; v < NVARS; v++)
            cons[v] = u[blk*u_bsz + (long long) p*LD + COFF(e, v, NVARS)];

        const fpdtype_t invrho = FP(1.0)/cons[0];
        fpdtype_t ke = 0;
        pri[0] = cons[0];
        UNROLL for (int i = 0; i < NDIMS; i++)
        {{
            pri[i + 1] = invrho*cons[i + 1];
            ke += cons[i + 1]*cons[i + 1];
        }}
        pri[NVARS - 1] = C_GM1*(cons[NVARS - 1] - FP(0.5)*invrho*ke);
{gsrc}{psrc}
        {wsrc}
{esrc}
    }}

    UNROLL for (int j = 0; j < NEXPRS; j++)
        out[blk*out_bsz + (long long) j*out_ld + e] = acc[j];
}}
'''
    return src, 'fieldeval', [a.split()[-1].lstrip('*') for a in args]


def wavespeed_source(be, tplargs, npts, pts):
    """``wavespeed`` (``pyfr/solvers/euler/kernels/wavespeed.mako``): per
    element, the largest over its solution points of
    ``sum_i |(S_i/|J|).v| + c |S_i/|J||`` -- the quantity the CFL
    controller divides by.  One thread owns one element and walks its
    points (the ``reduce(max)`` of the reference's kernel spec), so the
    per-element result needs no atomics."""
    nd, nv = tplargs['ndims'], tplargs['nvars']

    defs = [('NDIMS', nd), ('NVARS', nv), ('NPTS', npts),
            ('NVERTS', tplargs.get('nverts', 0)), ('NEED_RCPDJAC', 1)]
    defs += ph.physics_defines(tplargs['c'])

    gsrc, gargs, gbody = _geom(tplargs, pts)

    args = (['int neles', 'const fpdtype_t* __restrict__ u',
             'long long u_bsz', 'fpdtype_t* __restrict__ wspd',
             'long long wspd_bsz'] + gargs)

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz, defs)}
#define LD (NVARS*C_SUB)
{ph.flux_src}
{ph.geom_src}
{gsrc}

extern "C" __global__ void __launch_bounds__(128)
wavespeed({', '.join(args)})
{{
    const long long gid = (long long) blockIdx.x*blockDim.x + threadIdx.x;
    const int e = (int) (gid % C_SUB);
    const long long blk = gid / C_SUB;

    if (blk*C_SUB + e >= neles)
        return;

    fpdtype_t wmax = FP(0.0);

    for (int p = 0; p < NPTS; p++)
    {{
{gbody}
        fpdtype_t us[NVARS], ft[NDIMS][NVARS], pr, vel[NDIMS];
        UNROLL for (int v = 0; v < NVARS; v++)
            us[v] = u[blk*u_bsz + (long long) p*LD + COFF(e, v, NVARS)];

        inviscid_flux(us, ft, pr, vel);
        const fpdtype_t csnd = sqrt(C_GAMMA*pr/us[0]);

        fpdtype_t lam = FP(0.0);
        UNROLL for (int i = 0; i < NDIMS; i++)
        {{
            fpdtype_t sv = FP(0.0), ss = FP(0.0);
            UNROLL for (int j = 0; j < NDIMS; j++)
            {{
                const fpdtype_t sij = s[i][j]*rcpdjac_v;
                sv += sij*vel[j];
                ss += sij*sij;
            }}
            lam += fabs(sv) + csnd*sqrt(ss);
        }}

        wmax = fmax(wmax, lam);
    }}

    wspd[blk*wspd_bsz + e] = wmax;
}}
'''
    return src, 'wavespeed', [a.split()[-1].lstrip('*') for a in args]
