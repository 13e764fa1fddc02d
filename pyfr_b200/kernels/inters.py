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


_pair_src = r'''
// Two consecutive interface points per thread.  Points are sorted by their
// left-hand address, so the lanes of a block that share a flux-point row
// follow one another: where a side's two addresses are adjacent and the
// first is 16-byte aligned the pair moves with one 128-bit access.
struct pair_t
{
    ixdtype_t a0, a1;
    bool vec;
};

static __device__ __forceinline__ pair_t
pair_of(const fpdtype_t *base, const ixdtype_t *__restrict__ map,
        ixdtype_t i, bool two)
{
    pair_t p;
    p.a0 = map[i];
    p.a1 = two ? map[i + 1] : p.a0;
    p.vec = two && p.a1 == p.a0 + 1 &&
            (reinterpret_cast<unsigned long long>(base + p.a0)
             & (2*sizeof(fpdtype_t) - 1)) == 0;
    return p;
}

static __device__ __forceinline__ void
ld_pair(const fpdtype_t *__restrict__ b, const pair_t &p, fpdtype_t &x0,
        fpdtype_t &x1)
{
    if (p.vec)
    {
        const fpdtype2_t t = *reinterpret_cast<const fpdtype2_t *>(b + p.a0);
        x0 = t.x; x1 = t.y;
    }
    else
    {
        x0 = b[p.a0]; x1 = b[p.a1];
    }
}

static __device__ __forceinline__ void
st_pair(fpdtype_t *__restrict__ b, const pair_t &p, bool two, fpdtype_t x0,
        fpdtype_t x1)
{
    if (p.vec)
    {
        fpdtype2_t t;
        t.x = x0; t.y = x1;
        *reinterpret_cast<fpdtype2_t *>(b + p.a0) = t;
    }
    else
    {
        b[p.a0] = x0;
        if (two)
            b[p.a1] = x1;
    }
}
'''


def conu_pairs_source(be, tplargs, both=False):
    """Interior ``intconu`` with two consecutive points per thread (backend
    option ``conu-pairs``); same arguments as ``conu_source``, launched
    over ``ceil(n/2)`` threads."""
    nv, beta = tplargs['nvars'], tplargs['c']['ldg-beta']

    args = (['ixdtype_t n'] + _view_arg('ulin') + _view_arg('urin') +
            _view_arg('ulout', const=False) + _view_arg('urout', const=False))

    need_l, need_r = beta != 0.5, beta != -0.5
    st_l = both or beta != -0.5
    st_r = both or beta != 0.5

    if beta == -0.5:
        com = ('l0', 'l1')
    elif beta == 0.5:
        com = ('r0', 'r1')
    else:
        com = tuple(f'r{k}*{ph.fpconst(0.5 + beta)} + '
                    f'l{k}*{ph.fpconst(0.5 - beta)}' for k in (0, 1))

    L = []
    if need_l:
        L.append('const pair_t pli = pair_of(ulin, ulin_map, i, two);')
    if need_r:
        L.append('const pair_t pri = pair_of(urin, urin_map, i, two);')
    if st_l:
        L.append('const pair_t plo = pair_of(ulout, ulout_map, i, two);')
    if st_r:
        L.append('const pair_t pro = pair_of(urout, urout_map, i, two);')

    B = []
    if need_l:
        B.append('fpdtype_t l0, l1; ld_pair(ulin + K_SOA*v, pli, l0, l1);')
    if need_r:
        B.append('fpdtype_t r0, r1; ld_pair(urin + K_SOA*v, pri, r0, r1);')
    B.append(f'const fpdtype_t c0 = {com[0]}, c1 = {com[1]};')
    if st_l:
        B.append('st_pair(ulout + K_SOA*v, plo, two, c0, c1);')
    if st_r:
        B.append('st_pair(urout + K_SOA*v, pro, two, c0, c1);')

    nl = '\n    '
    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz, [('NVARS', nv)])}
{_pair_src}
extern "C" __global__ void __launch_bounds__(128)
intconu({', '.join(args)})
{{
    const ixdtype_t i = 2*((ixdtype_t) blockIdx.x*blockDim.x + threadIdx.x);
    if (i >= n)
        return;
    const bool two = i + 1 < n;

    {nl.join(L)}

    UNROLL for (int v = 0; v < NVARS; v++)
    {{
        {(nl + '    ').join(B)}
    }}
}}
'''
    return src, 'intconu', _names(args)


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


# -- boundary conditions --------------------------------------------------------
def _bc_states_src(tplargs):
    """Device functions ``bc_rsolve_state`` / ``bc_ldg_state`` /
    ``bc_ldg_grad_state`` of one boundary type
    (``pyfr/solvers/{euler,navstokes}/kernels/bcs/<type>.mako``).  Section
    parameters arrive as C expressions in ``ploc[i]`` and ``t`` (built by
    the host exactly like the reference's ``_exp_opts``) or as numbers."""
    nd, nv, c = tplargs['ndims'], tplargs['nvars'], tplargs['c']
    bt = tplargs['bctype']
    uvw = 'uvw'[:nd]

    def val(k):
        v = c[k]
        return f'FP({v})' if isinstance(v, str) else ph.fpconst(v)

    ke = lambda a: ' + '.join(f'{a}[{i + 1}]*{a}[{i + 1}]' for i in range(nd))
    sig = ('(const fpdtype_t ul[NVARS], const fpdtype_t nl[NDIMS], '
           'fpdtype_t ur[NVARS], const fpdtype_t ploc[NDIMS], '
           'const fpdtype_t t)')
    L = []

    def fn(name, body):
        L.append(f'__device__ __forceinline__ void {name}{sig}\n{{\n'
                 + '\n'.join('    ' + b for b in body) + '\n}\n')

    def alias(name, target):
        fn(name, [f'{target}(ul, nl, ur, ploc, t);'])

    if bt == 'no-slp-adia-wall':
        fn('bc_rsolve_state',
           ['ur[0] = ul[0];'] + [f'ur[{i + 1}] = -ul[{i + 1}];'
                                 for i in range(nd)]
           + ['ur[NVARS - 1] = ul[NVARS - 1];'])
        fn('bc_ldg_state',
           ['ur[0] = ul[0];'] + [f'ur[{i + 1}] = FP(0.0);' for i in range(nd)]
           + [f'ur[NVARS - 1] = ul[NVARS - 1] - (FP(0.5)/ul[0])*({ke("ul")});'])
    elif bt == 'slp-adia-wall':
        nor = ' + '.join(f'ul[{i + 1}]*nl[{i}]' for i in range(nd))
        fn('bc_rsolve_state',
           [f'const fpdtype_t nor = {nor};', 'ur[0] = ul[0];']
           + [f'ur[{i + 1}] = ul[{i + 1}] - FP(2.0)*nor*nl[{i}];'
              for i in range(nd)] + ['ur[NVARS - 1] = ul[NVARS - 1];'])
        alias('bc_ldg_state', 'bc_rsolve_state')
    elif bt == 'no-slp-isot-wall':
        ct = ph.fpconst(c['cpTw']/c['gamma'])
        tail = [f'ur[NVARS - 1] = {ct}*ur[0] + FP(0.5)*(FP(1.0)/ur[0])*'
                f'({ke("ur")});']
        fn('bc_rsolve_state',
           ['ur[0] = ul[0];']
           + [f'ur[{i + 1}] = -ul[{i + 1}] + FP(2.0)*{val(v)}*ul[0];'
              for i, v in enumerate(uvw)] + tail)
        fn('bc_ldg_state',
           ['ur[0] = ul[0];'] + [f'ur[{i + 1}] = {val(v)}*ul[0];'
                                 for i, v in enumerate(uvw)] + tail)
    elif bt == 'sup-out-fn':
        fn('bc_rsolve_state', ['UNROLL for (int i = 0; i < NVARS; i++)',
                               '    ur[i] = ul[i];'])
        alias('bc_ldg_state', 'bc_rsolve_state')
    elif bt in ('sup-in-fa', 'sub-in-frv'):
        body = [f'ur[0] = {val("rho")};'] + [
            f'ur[{i + 1}] = ({val("rho")})*({val(v)});'
            for i, v in enumerate(uvw)]
        if bt == 'sup-in-fa':
            body.append(f'ur[NVARS - 1] = {val("p")}/C_GM1 + FP(0.5)*'
                        f'(FP(1.0)/ur[0])*({ke("ur")});')
        else:
            body.append('ur[NVARS - 1] = ul[NVARS - 1] - FP(0.5)*'
                        f'(FP(1.0)/ul[0])*({ke("ul")}) + FP(0.5)*'
                        f'(FP(1.0)/ur[0])*({ke("ur")});')
        fn('bc_rsolve_state', body)
        alias('bc_ldg_state', 'bc_rsolve_state')
    elif bt == 'sub-out-fp':
        fn('bc_rsolve_state',
           ['UNROLL for (int i = 0; i < NVARS - 1; i++)', '    ur[i] = ul[i];',
            f'ur[NVARS - 1] = {val("p")}/C_GM1 + FP(0.5)*(FP(1.0)/ul[0])*'
            f'({ke("ul")});'])
        alias('bc_ldg_state', 'bc_rsolve_state')
    elif bt == 'sub-in-ftpttang':
        rdcp = c['Rdcp']
        body = [
            f'const fpdtype_t pl = C_GM1*(ul[NVARS - 1] - (FP(0.5)/ul[0])*'
            f'({ke("ul")}));',
            f'fpdtype_t udotu = {ph.fpconst(2.0*c["cpTt"])}*(FP(1.0) - '
            f'{ph.fpconst(c["pt"]**(-rdcp))}*pow(pl, {ph.fpconst(rdcp)}));',
            'udotu = fmax(FP(0.0), udotu);',
            f'ur[0] = {ph.fpconst(1.0/rdcp)}*pl/({ph.fpconst(c["cpTt"])} - '
            'FP(0.5)*udotu);'
        ]
        body += [f'ur[{i + 1}] = {ph.fpconst(v)}*ur[0]*sqrt(udotu);'
                 for i, v in enumerate(c['vc'])]
        body.append('ur[NVARS - 1] = (FP(1.0)/C_GM1)*pl + '
                    'FP(0.5)*ur[0]*udotu;')
        fn('bc_rsolve_state', body)
        alias('bc_ldg_state', 'bc_rsolve_state')
    elif bt == 'char-riem-inv':
        Ve = ' + '.join(f'({val(v)})*nl[{i}]' for i, v in enumerate(uvw))
        Vi = ' + '.join(f'ul[{i + 1}]*nl[{i}]' for i in range(nd))
        body = [
            f'const fpdtype_t pe = {val("p")}, rhoe = {val("rho")};',
            'const fpdtype_t cs = sqrt(C_GAMMA*pe/rhoe);',
            'const fpdtype_t s = pe*pow(rhoe, -C_GAMMA);',
            'const fpdtype_t ratio = cs*(FP(2.0)/C_GM1);',
            'const fpdtype_t inv = FP(1.0)/ul[0];',
            f'const fpdtype_t V_e = {Ve};',
            f'const fpdtype_t V_i = inv*({Vi});',
            'const fpdtype_t p_i = C_GM1*ul[NVARS - 1] - (FP(0.5)*C_GM1)*inv*'
            f'({ke("ul")});',
            'const fpdtype_t c_i = sqrt(C_GAMMA*p_i*inv);',
            'const fpdtype_t R_e = (fabs(V_e) >= cs && V_i >= 0) '
            '? V_i - c_i*(FP(2.0)/C_GM1) : V_e - ratio;',
            'const fpdtype_t R_i = (fabs(V_e) >= cs && V_i < 0) '
            '? V_e + ratio : V_i + c_i*(FP(2.0)/C_GM1);',
            'const fpdtype_t V_b = FP(0.5)*(R_e + R_i);',
            'const fpdtype_t c_b = (FP(0.25)*C_GM1)*(R_i - R_e);',
            'const fpdtype_t rho_b = (V_i < 0) '
            '? pow((FP(1.0)/(C_GAMMA*s))*c_b*c_b, FP(1.0)/C_GM1) '
            ': ul[0]*pow(ul[0]*c_b*c_b/(C_GAMMA*p_i), FP(1.0)/C_GM1);',
            'const fpdtype_t p_b = (FP(1.0)/C_GAMMA)*rho_b*c_b*c_b;',
            'ur[0] = rho_b;'
        ]
        body += [f'ur[{i + 1}] = (V_i >= 0) '
                 f'? rho_b*(ul[{i + 1}]*inv + (V_b - V_i)*nl[{i}]) '
                 f': rho_b*(({val(v)}) + (V_b - V_e)*nl[{i}]);'
                 for i, v in enumerate(uvw)]
        body.append('ur[NVARS - 1] = p_b*(FP(1.0)/C_GM1) + FP(0.5)*'
                    f'(FP(1.0)/ur[0])*({ke("ur")});')
        fn('bc_rsolve_state', body)
        alias('bc_ldg_state', 'bc_rsolve_state')
    else:
        raise NotImplementedError(f'Boundary type {bt!r} is not on the b200 '
                                  'path (see DESIGN.md)')

    # Ghost-side gradient
    gsig = ('(const fpdtype_t ur[NVARS], const fpdtype_t nl[NDIMS], '
            'const fpdtype_t gul[NDIMS][NVARS], fpdtype_t gur[NDIMS][NVARS])')
    if bt in ('char-riem-inv', 'sup-in-fa', 'sub-in-frv', 'sub-out-fp'):
        gbody = ['UNROLL for (int d = 0; d < NDIMS; d++)',
                 '    UNROLL for (int v = 0; v < NVARS; v++)',
                 '        gur[d][v] = FP(0.0);']
    elif bt == 'no-slp-adia-wall':
        # remove the wall-normal temperature gradient
        # (no-slp-adia-wall.mako:22-75)
        gbody = [
            'const fpdtype_t rcprho = FP(1.0)/ur[0];',
            'fpdtype_t vel[NDIMS], Tl[NDIMS];',
            'UNROLL for (int i = 0; i < NDIMS; i++)',
            '    vel[i] = rcprho*ur[i + 1];',
            'UNROLL for (int d = 0; d < NDIMS; d++)', '{',
            '    fpdtype_t acc = rcprho*gul[d][0]*ur[NVARS - 1];',
            '    UNROLL for (int i = 0; i < NDIMS; i++)',
            '        acc += vel[i]*(gul[d][i + 1] - vel[i]*gul[d][0]);',
            '    Tl[d] = gul[d][NVARS - 1] - acc;', '}',
            'UNROLL for (int d = 0; d < NDIMS; d++)',
            '    UNROLL for (int v = 0; v < NVARS; v++)',
            '        gur[d][v] = gul[d][v];',
            'UNROLL for (int d = 0; d < NDIMS; d++)', '{',
            '    fpdtype_t acc = 0;',
            '    UNROLL for (int e = 0; e < NDIMS; e++)',
            '        acc += nl[d]*nl[e]*Tl[e];',
            '    gur[d][NVARS - 1] -= acc;', '}']
    else:
        gbody = ['UNROLL for (int d = 0; d < NDIMS; d++)',
                 '    UNROLL for (int v = 0; v < NVARS; v++)',
                 '        gur[d][v] = gul[d][v];']

    L.append(f'__device__ __forceinline__ void bc_ldg_grad_state{gsig}\n{{\n'
             + '\n'.join('    ' + b for b in gbody) + '\n}\n')

    return '\n'.join(L)


def bc_source(be, tplargs, viscous, kind, has_ploc):
    """``bcconu`` (``pyfr/solvers/navstokes/kernels/bcconu.mako``) and
    ``bccflux`` (``pyfr/solvers/euler/kernels/bccflux.mako``,
    ``pyfr/solvers/navstokes/kernels/bccflux.mako`` with the flux-state
    templates ``bcs/ghost.mako`` / ``bcs/ghost-imperm.mako``)."""
    nd, nv, c = tplargs['ndims'], tplargs['nvars'], tplargs['c']
    cfs = tplargs.get('bccfluxstate')

    defs = [('NDIMS', nd), ('NVARS', nv),
            ('C_GM1', ph.fpconst(c['gamma'] - 1))]
    defs += ph.physics_defines(c, tplargs.get('visc_corr', 'none'), viscous)

    head = _head + '    const fpdtype_t t = *t_p;\n' + _normal_src()
    if has_ploc:
        head += r'''
    fpdtype_t ploc[NDIMS];
    UNROLL for (int d = 0; d < NDIMS; d++)
        ploc[d] = __ldg(plocp + (long long) d*ploc_ld + i);
'''
    else:
        head += '    const fpdtype_t ploc[NDIMS] = {};\n'

    tail_args = ['const fpdtype_t* __restrict__ nl', 'long long nl_ld']
    if has_ploc:
        tail_args += ['const fpdtype_t* __restrict__ plocp',
                      'long long ploc_ld']
    tail_args += ['const fpdtype_t* __restrict__ t_p']

    if kind == 'bcconu':
        args = (['ixdtype_t n'] + _view_arg('ulin')
                + _view_arg('ulout', const=False) + tail_args)
        body = head + r'''
    fpdtype_t l[NVARS], r[NVARS];
    UNROLL for (int v = 0; v < NVARS; v++)
        l[v] = ulin[ulin_map[i] + K_SOA*v];

    bc_ldg_state(l, nrm, r, ploc, t);

    UNROLL for (int v = 0; v < NVARS; v++)
        ulout[ulout_map[i] + K_SOA*v] = r[v];
'''
    else:
        args = ['ixdtype_t n'] + _view_arg('ul', const=False)
        if viscous:
            args += _view_arg('gradul', strided=True)
        args += tail_args

        body = head + r'''
    const ixdtype_t lix = ul_map[i];
    fpdtype_t l[NVARS], r[NVARS], fn[NVARS];
    UNROLL for (int v = 0; v < NVARS; v++)
        l[v] = ul[lix + K_SOA*v];
'''
        if viscous and cfs is not None:
            state = 'l' if cfs == 'ghost' else 'r'
            tau = c['ldg-tau'] if cfs == 'ghost' else 0.0
            body += rf'''
    fpdtype_t gl[NDIMS][NVARS], gr[NDIMS][NVARS], fvr[NDIMS][NVARS] = {{}};
    {{
        const ixdtype_t gix = gradul_map[i], gst = gradul_str[i];
        UNROLL for (int d = 0; d < NDIMS; d++)
            UNROLL for (int v = 0; v < NVARS; v++)
                gl[d][v] = gradul[gix + gst*d + K_SOA*v];
    }}

    // Viscous ghost state and flux
    bc_ldg_state(l, nrm, r, ploc, t);
    bc_ldg_grad_state({state}, nrm, gl, gr);
    viscous_flux_add(r, gr, fvr);

    // Inviscid ghost state and Riemann solve
    bc_rsolve_state(l, nrm, r, ploc, t);
    rsolve(l, r, nrm, fn);

    UNROLL for (int v = 0; v < NVARS; v++)
    {{
        fpdtype_t fv = {' + '.join(f'nrm[{j}]*fvr[{j}][v]'
                                   for j in range(nd))};
        {f'fv += {ph.fpconst(tau)}*(l[v] - r[v]);' if tau != 0.0 else ''}
        ul[lix + K_SOA*v] = mag_nl*(fn[v] + fv);
    }}
'''
        else:
            body += r'''
    bc_rsolve_state(l, nrm, r, ploc, t);
    rsolve(l, r, nrm, fn);

    UNROLL for (int v = 0; v < NVARS; v++)
        ul[lix + K_SOA*v] = mag_nl*fn[v];
'''

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz, defs)}
{ph.flux_src}
{ph.visc_src if viscous else ''}
{ph.rsolve_src[tplargs['rsolver']] if kind == 'bccflux' else ''}
{_bc_states_src(tplargs)}

extern "C" __global__ void __launch_bounds__(128)
{kind}({', '.join(args)})
{{
{body}
}}
'''
    return src, kind, _names(args)
