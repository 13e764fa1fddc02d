"""CUDA source for the interface kernels and the halo pack kernel.

``intcflux``/``mpicflux`` (Euler: ``pyfr/solvers/euler/kernels/{intcflux,
mpicflux}.mako``; Navier-Stokes: ``pyfr/solvers/navstokes/kernels/{intcflux,
mpicflux}.mako``), ``intconu``/``mpiconu`` (``pyfr/solvers/navstokes/
kernels/{intconu,mpiconu}.mako``) and ``pack`` (``pyfr/backends/base/
kernels/packing.mako``).

One thread per interface flux point.  A view argument is a base pointer
plus a per-point element offset (``*_map``) and, for gradients, a per-point
row stride (``*_str``); variable ``v`` sits ``K_SOA*v`` elements further on
(reference ``pyfr/backends/base/generator.py:171-202``).  Interior
interfaces are sorted by left-hand address at set-up, so left-side accesses
of a warp fall into whole row segments; right-side accesses are whatever
the mesh connectivity dictates.  Data received from a neighbouring
partition is a dense ``[nvars][n]`` / ``[ndims*nvars][n]`` matrix
(``:214-226``).  Each side of an interface is read once and, for the flux
kernels, written once.
"""

from pyfr_b200.kernels import physics as ph

_head = r'''
    const ixdtype_t i = (ixdtype_t) blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n)
        return;
'''


def _view_arg(name, const=True, strided=False):
    q = 'const ' if const else ''
    a = [f'{q}fpdtype_t* __restrict__ {name}',
         f'const ixdtype_t* __restrict__ {name}_map']
    if strided:
        a.append(f'const ixdtype_t* __restrict__ {name}_str')
    return a


def _mpi_arg(name):
    return [f'const fpdtype_t* __restrict__ {name}']


def _names(args):
    return [a.split()[-1].lstrip('*') for a in args]


def _normal_src():
    return r'''
    fpdtype_t nrm[NDIMS], mag2 = 0;
    UNROLL for (int d = 0; d < NDIMS; d++)
    {
        nrm[d] = __ldg(nl + (long long) d*nl_ld + i);
        mag2 += nrm[d]*nrm[d];
    }

    const fpdtype_t mag_nl = sqrt(mag2), rcpmag = FP(1.0)/mag_nl;
    UNROLL for (int d = 0; d < NDIMS; d++)
        nrm[d] *= rcpmag;
'''


def cflux_source(be, tplargs, viscous, mpi):
    nd, nv, c = tplargs['ndims'], tplargs['nvars'], tplargs['c']
    name = 'mpicflux' if mpi else 'intcflux'

    defs = [('NDIMS', nd), ('NVARS', nv)]
    defs += ph.physics_defines(c, tplargs.get('visc_corr', 'none'), viscous)

    args = ['ixdtype_t n'] + _view_arg('ul', const=False)
    args += _mpi_arg('ur') if mpi else _view_arg('ur', const=False)

    body = r'''
    const ixdtype_t lix = ul_map[i];
    fpdtype_t l[NVARS], r[NVARS];
    UNROLL for (int v = 0; v < NVARS; v++)
        l[v] = ul[lix + K_SOA*v];
'''
    if mpi:
        body += r'''
    UNROLL for (int v = 0; v < NVARS; v++)
        r[v] = __ldg(ur + (long long) v*n + i);
'''
    else:
        body += r'''
    const ixdtype_t rix = ur_map[i];
    UNROLL for (int v = 0; v < NVARS; v++)
        r[v] = ur[rix + K_SOA*v];
'''

    body += _normal_src()
    body += r'''
    fpdtype_t fn[NVARS];
    rsolve(l, r, nrm, fn);
'''

    if viscous:
        beta, tau = c['ldg-beta'], c['ldg-tau']
        defs += [('C_TAU', ph.fpconst(tau))]
        need_l, need_r = beta != -0.5, beta != 0.5

        if need_l:
            args += _view_arg('gradul', strided=True)
            body += r'''
    fpdtype_t gl[NDIMS][NVARS], fvl[NDIMS][NVARS] = {};
    {
        const ixdtype_t gix = gradul_map[i], gst = gradul_str[i];
        UNROLL for (int d = 0; d < NDIMS; d++)
            UNROLL for (int v = 0; v < NVARS; v++)
                gl[d][v] = gradul[gix + gst*d + K_SOA*v];
    }
    viscous_flux_add(l, gl, fvl);
'''
        if need_r:
            if mpi:
                args += _mpi_arg('gradur')
                body += r'''
    fpdtype_t gr[NDIMS][NVARS], fvr[NDIMS][NVARS] = {};
    UNROLL for (int d = 0; d < NDIMS; d++)
        UNROLL for (int v = 0; v < NVARS; v++)
            gr[d][v] = __ldg(gradur + (long long) (NVARS*d + v)*n + i);
    viscous_flux_add(r, gr, fvr);
'''
            else:
                args += _view_arg('gradur', strided=True)
                body += r'''
    fpdtype_t gr[NDIMS][NVARS], fvr[NDIMS][NVARS] = {};
    {
        const ixdtype_t gix = gradur_map[i], gst = gradur_str[i];
        UNROLL for (int d = 0; d < NDIMS; d++)
            UNROLL for (int v = 0; v < NVARS; v++)
                gr[d][v] = gradur[gix + gst*d + K_SOA*v];
    }
    viscous_flux_add(r, gr, fvr);
'''

        def ndot(f):
            return ' + '.join(f'nrm[{j}]*{f}[{j}][v]' for j in range(nd))

        if beta == -0.5:
            fv = ndot('fvr')
        elif beta == 0.5:
            fv = ndot('fvl')
        else:
            fv = (f'{ph.fpconst(0.5 + beta)}*({ndot("fvl")}) + '
                  f'{ph.fpconst(0.5 - beta)}*({ndot("fvr")})')

        body += f'''
    UNROLL for (int v = 0; v < NVARS; v++)
    {{
        fpdtype_t fv = {fv};
        {'fv += C_TAU*(l[v] - r[v]);' if tau != 0.0 else ''}
        fn[v] += fv;
    }}
'''

    body += r'''
    UNROLL for (int v = 0; v < NVARS; v++)
    {
        const fpdtype_t fc = mag_nl*fn[v];
        ul[lix + K_SOA*v] = fc;
'''
    if not mpi:
        body += '        ur[rix + K_SOA*v] = -fc;\n'
    body += '    }\n'

    args += ['const fpdtype_t* __restrict__ nl', 'long long nl_ld']

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz, defs)}
{ph.flux_src}
{ph.visc_src if viscous else ''}
{ph.rsolve_src[tplargs['rsolver']]}

extern "C" __global__ void
__launch_bounds__(128, {getattr(be, 'cflux_minblocks', 4) if viscous else 8})
{name}({', '.join(args)})
{{
{_head}
{body}
}}
'''
    return src, name, _names(args)


def conu_source(be, tplargs, mpi, both=False):
    """``both``: also store each side's own trace where the reference
    kernel leaves it untouched (|beta| = 1/2), which makes the preceding
    ``copy_fpts`` pass over the whole flux-point array redundant."""
    nv, beta = tplargs['nvars'], tplargs['c']['ldg-beta']
    name = 'mpiconu' if mpi else 'intconu'

    args = ['ixdtype_t n'] + _view_arg('ulin')
    args += _mpi_arg('urin') if mpi else _view_arg('urin')
    args += _view_arg('ulout', const=False)
    if not mpi:
        args += _view_arg('urout', const=False)

    ldl = 'ulin[ulin_map[i] + K_SOA*v]'
    ldr = ('__ldg(urin + (long long) v*n + i)' if mpi
           else 'urin[urin_map[i] + K_SOA*v]')

    if mpi:
        if beta == -0.5:
            stmt = f'ulout[ulout_map[i] + K_SOA*v] = {ldl};'
        elif beta == 0.5:
            stmt = f'ulout[ulout_map[i] + K_SOA*v] = {ldr};'
        else:
            stmt = (f'ulout[ulout_map[i] + K_SOA*v] = '
                    f'{ldr}*{ph.fpconst(0.5 + beta)} + '
                    f'{ldl}*{ph.fpconst(0.5 - beta)};')
    else:
        if beta == -0.5 and both:
            stmt = (f'const fpdtype_t com = {ldl};\n'
                    '        ulout[ulout_map[i] + K_SOA*v] = com;\n'
                    '        urout[urout_map[i] + K_SOA*v] = com;')
        elif beta == 0.5 and both:
            stmt = (f'const fpdtype_t com = {ldr};\n'
                    '        ulout[ulout_map[i] + K_SOA*v] = com;\n'
                    '        urout[urout_map[i] + K_SOA*v] = com;')
        elif beta == -0.5:
            stmt = f'urout[urout_map[i] + K_SOA*v] = {ldl};'
        elif beta == 0.5:
            stmt = f'ulout[ulout_map[i] + K_SOA*v] = {ldr};'
        else:
            stmt = (f'const fpdtype_t com = {ldr}*{ph.fpconst(0.5 + beta)} + '
                    f'{ldl}*{ph.fpconst(0.5 - beta)};\n'
                    '        ulout[ulout_map[i] + K_SOA*v] = com;\n'
                    '        urout[urout_map[i] + K_SOA*v] = com;')

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz, [('NVARS', nv)])}

extern "C" __global__ void __launch_bounds__(128)
{name}({', '.join(args)})
{{
{_head}
    UNROLL for (int v = 0; v < NVARS; v++)
    {{
        {stmt}
    }}
}}
'''
    return src, name, _names(args)


def pack_source(be, nrv, ncv):
    args = (['ixdtype_t n'] + _view_arg('v', strided=nrv > 1) +
            ['fpdtype_t* __restrict__ pmat'])

    rs = 'v_str[i]*r + ' if nrv > 1 else ''

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz,
                          [('NRV', nrv), ('NCV', ncv)])}

extern "C" __global__ void __launch_bounds__(128)
pack_view({', '.join(args)})
{{
{_head}
    const ixdtype_t ix = v_map[i];

    UNROLL for (int r = 0; r < NRV; r++)
        UNROLL for (int c = 0; c < NCV; c++)
            pmat[(long long) (r*NCV + c)*n + i] = v[ix + {rs}K_SOA*c];
}}
'''
    return src, 'pack_view', _names(args)
