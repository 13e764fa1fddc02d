"""CUDA source of the device functions shared by the flux kernels.

Hand-written sm_100a device code for the arithmetic the reference keeps in
Mako macros: ``inviscid_flux`` (``pyfr/solvers/euler/kernels/flux.mako``),
``viscous_flux_add`` (``pyfr/solvers/navstokes/kernels/flux.mako``),
the Rusanov and HLLC Riemann solvers (``pyfr/solvers/euler/kernels/
rsolvers/{rusanov,hllc}.mako``), ``transform_grad`` (``pyfr/solvers/
baseadvecdiff/kernels/transform_grad.mako``) and the linear-element metric
terms (``pyfr/solvers/baseadvec/kernels/smats.mako``).  Everything is
dimension-generic over the compile-time constants ``NDIMS``/``NVARS`` and
lives in registers after full unrolling.
"""


def prologue(fpdtype, ixdtype, soasz, csubsz, defines=()):
    """Common header: scalar types, layout constants, addressing helpers."""
    fp = {'float64': 'double', 'float32': 'float'}[str(fpdtype)]
    ix = {'int32': 'int', 'int64': 'long long'}[str(ixdtype)]

    lines = [
        f'typedef {fp} fpdtype_t;',
        f'typedef {ix} ixdtype_t;',
        f'typedef {fp}2 fpdtype2_t;',
        f'#define PYFR_B200_FP64 {int(fp == "double")}',
        f'#define K_SOA {soasz}',
        f'#define C_SUB {csubsz}',
        *[f'#define {k} {v}' for k, v in defines],
        # Offset of (variable v, element e of the block) inside one row of a
        # stacked matrix holding nv variables
        '#define COFF(e, v, nv) (((e)/K_SOA)*(K_SOA*(nv)) + (v)*K_SOA'
        ' + (e) % K_SOA)',
        '#define FP(x) ((fpdtype_t) (x))',
        '#define UNROLL _Pragma("unroll")',
        '// optimisation barrier: the value is re-formed where it is used',
        '#define OPAQUE(x) asm volatile("" : "+r"(x))',
        '#ifdef __CUDACC__',
        '#define OPAQUE64(x) asm volatile("" : "+l"(x))',
        '#else',
        '#define OPAQUE64(x) asm volatile("" : "+r"(x))',
        '#endif',
        ''
    ]

    return '\n'.join(lines)


def fpconst(v):
    """Exact C literal of a Python float, in the working precision."""
    return f'FP({float(v)!r})'


flux_src = r'''
// ---- inviscid (Euler) flux: f[d][v], pressure, velocity -------------------
__device__ __forceinline__ void
inviscid_flux(const fpdtype_t s[NVARS], fpdtype_t f[NDIMS][NVARS],
              fpdtype_t &p, fpdtype_t v[NDIMS])
{
    const fpdtype_t invrho = FP(1.0)/s[0], E = s[NVARS - 1];
    fpdtype_t rhov[NDIMS], ke = 0;

    UNROLL for (int i = 0; i < NDIMS; i++)
    {
        rhov[i] = s[i + 1];
        v[i] = invrho*rhov[i];
        ke += rhov[i]*rhov[i];
    }

    p = (C_GAMMA - FP(1.0))*(E - FP(0.5)*invrho*ke);

    UNROLL for (int i = 0; i < NDIMS; i++)
    {
        f[i][0] = rhov[i];
        f[i][NVARS - 1] = (E + p)*v[i];

        UNROLL for (int j = 0; j < NDIMS; j++)
            f[i][j + 1] = rhov[i]*v[j] + ((i == j) ? p : FP(0.0));
    }
}
'''

visc_src = r'''
// ---- viscous flux, added into fout[d][v] -----------------------------------
// The flux is linear in the gradient: ``sc`` scales it through the
// viscosity, so a caller holding the gradient without its 1/|J| factor
// passes that factor here instead of multiplying NDIMS*NVARS values by it.
__device__ __forceinline__ void
viscous_flux_add_sc(const fpdtype_t u[NVARS],
                    const fpdtype_t g[NDIMS][NVARS],
                    fpdtype_t fout[NDIMS][NVARS], const fpdtype_t sc)
{
    const fpdtype_t rho = u[0], E = u[NVARS - 1];
    const fpdtype_t rcprho = FP(1.0)/rho;
    fpdtype_t vel[NDIMS], dv[NDIMS][NDIMS], T_x[NDIMS], t[NDIMS][NDIMS];

    UNROLL for (int i = 0; i < NDIMS; i++)
        vel[i] = rcprho*u[i + 1];

    // dv[i][d] = rho * d(vel_i)/d(x_d)
    UNROLL for (int i = 0; i < NDIMS; i++)
        UNROLL for (int d = 0; d < NDIMS; d++)
            dv[i][d] = g[d][i + 1] - vel[i]*g[d][0];

#ifdef VISC_SUTHERLAND
    fpdtype_t q2 = 0;
    UNROLL for (int i = 0; i < NDIMS; i++)
        q2 += vel[i]*vel[i];
    const fpdtype_t cpT = C_GAMMA*(rcprho*E - FP(0.5)*q2);
    const fpdtype_t Trat = C_RCPCPTREF*cpT;
    const fpdtype_t mu_c = sc*(C_MU_SUTH*Trat*sqrt(Trat)/(cpT + C_CPTS));
#else
    const fpdtype_t mu_c = sc*C_MU;
#endif

    fpdtype_t div = 0;
    UNROLL for (int d = 0; d < NDIMS; d++)
    {
        fpdtype_t acc = rcprho*g[d][0]*E;
        UNROLL for (int i = 0; i < NDIMS; i++)
            acc += vel[i]*dv[i][d];

        T_x[d] = rcprho*(g[d][NVARS - 1] - acc);
        div += dv[d][d];
    }

    UNROLL for (int i = 0; i < NDIMS; i++)
    {
        t[i][i] = FP(-2.0)*mu_c*rcprho*(dv[i][i] - FP(1.0/3.0)*div);

        UNROLL for (int j = i + 1; j < NDIMS; j++)
            t[i][j] = t[j][i] = -mu_c*rcprho*(dv[j][i] + dv[i][j]);
    }

    UNROLL for (int d = 0; d < NDIMS; d++)
    {
        fpdtype_t ef = 0;

        UNROLL for (int i = 0; i < NDIMS; i++)
        {
            fout[d][i + 1] += t[d][i];
            ef += vel[i]*t[d][i];
        }

        fout[d][NVARS - 1] += ef + -mu_c*C_GAMMA_PR*T_x[d];
    }
}

__device__ __forceinline__ void
viscous_flux_add(const fpdtype_t u[NVARS], const fpdtype_t g[NDIMS][NVARS],
                 fpdtype_t fout[NDIMS][NVARS])
{
    viscous_flux_add_sc(u, g, fout, FP(1.0));
}
'''

rsolve_src = {
    'rusanov': r'''
// ---- Rusanov (local Lax-Friedrichs) ---------------------------------------
__device__ __forceinline__ void
rsolve(const fpdtype_t ul[NVARS], const fpdtype_t ur[NVARS],
       const fpdtype_t n[NDIMS], fpdtype_t nf[NVARS])
{
    fpdtype_t fl[NDIMS][NVARS], fr[NDIMS][NVARS], vl[NDIMS], vr[NDIMS];
    fpdtype_t pl, pr, nv = 0;

    inviscid_flux(ul, fl, pl, vl);
    inviscid_flux(ur, fr, pr, vr);

    UNROLL for (int i = 0; i < NDIMS; i++)
        nv += n[i]*(vl[i] + vr[i]);

    const fpdtype_t a = sqrt(FP(0.25)*C_GAMMA*(pl + pr)/(ul[0] + ur[0]))
                      + FP(0.25)*fabs(nv);

    UNROLL for (int i = 0; i < NVARS; i++)
    {
        fpdtype_t s = 0;
        UNROLL for (int j = 0; j < NDIMS; j++)
            s += n[j]*(fl[j][i] + fr[j][i]);

        nf[i] = FP(0.5)*s + a*(ul[i] - ur[i]);
    }
}
''',
    'hllc': r'''
// ---- HLLC -------------------------------------------------------------------
__device__ __forceinline__ void
rsolve(const fpdtype_t ul[NVARS], const fpdtype_t ur[NVARS],
       const fpdtype_t n[NDIMS], fpdtype_t nf[NVARS])
{
    fpdtype_t fl[NDIMS][NVARS], fr[NDIMS][NVARS], vl[NDIMS], vr[NDIMS];
    fpdtype_t va[NDIMS], usl[NVARS], usr[NVARS];
    fpdtype_t pl, pr, nvl = 0, nvr = 0, qq = 0;

    inviscid_flux(ul, fl, pl, vl);
    inviscid_flux(ur, fr, pr, vr);

    UNROLL for (int i = 0; i < NDIMS; i++)
    {
        nvl += n[i]*vl[i];
        nvr += n[i]*vr[i];
    }

    const fpdtype_t al = sqrt(C_GAMMA*pl/ul[0]), ar = sqrt(C_GAMMA*pr/ur[0]);
    const fpdtype_t srl = sqrt(ul[0]), srr = sqrt(ur[0]);

    // Roe averages
    const fpdtype_t nv = (srl*nvl + srr*nvr)/(srl + srr);
    const fpdtype_t H = (srl*(pr + ur[NDIMS + 1]) + srr*(pl + ul[NDIMS + 1]))
                      / (srl*ur[0] + srr*ul[0]);
    const fpdtype_t inv_rar = FP(1.0)/(srl + srr);

    UNROLL for (int i = 0; i < NDIMS; i++)
    {
        va[i] = (vl[i]*srl + vr[i]*srr)*inv_rar;
        qq += va[i]*va[i];
    }

    const fpdtype_t a = sqrt((C_GAMMA - FP(1.0))*(H - FP(0.5)*qq));

    // Wave speed estimates
    const fpdtype_t sl = fmin(nv - a, nvl - al);
    const fpdtype_t sr = fmax(nv + a, nvr + ar);
    const fpdtype_t sstar = (pr - pl + ul[0]*nvl*(sl - nvl)
                                     - ur[0]*nvr*(sr - nvr))
                          / (ul[0]*(sl - nvl) - ur[0]*(sr - nvr));

    const fpdtype_t ul_com = (sl - nvl)/(sl - sstar);
    const fpdtype_t ur_com = (sr - nvr)/(sr - sstar);

    usl[0] = ul_com*ul[0];
    usr[0] = ur_com*ur[0];

    UNROLL for (int i = 0; i < NDIMS; i++)
    {
        usl[i + 1] = usl[0]*(vl[i] + (sstar - nvl)*n[i]);
        usr[i + 1] = usr[0]*(vr[i] + (sstar - nvr)*n[i]);
    }

    usl[NVARS - 1] = ul_com*(ul[NVARS - 1] + (sstar - nvl)*
                             (ul[0]*sstar + pl/(sl - nvl)));
    usr[NVARS - 1] = ur_com*(ur[NVARS - 1] + (sstar - nvr)*
                             (ur[0]*sstar + pr/(sr - nvr)));

    UNROLL for (int i = 0; i < NVARS; i++)
    {
        fpdtype_t nf_fl = 0, nf_fr = 0;

        UNROLL for (int j = 0; j < NDIMS; j++)
        {
            nf_fl += n[j]*fl[j][i];
            nf_fr += n[j]*fr[j][i];
        }

        const fpdtype_t nf_fsl = nf_fl + sl*(usl[i] - ul[i]);
        const fpdtype_t nf_fsr = nf_fr + sr*(usr[i] - ur[i]);

        nf[i] = (0 <= sl) ? nf_fl : (sl <= 0 && 0 <= sstar) ? nf_fsl :
                (sstar <= 0 && 0 <= sr) ? nf_fsr : nf_fr;
    }
}
'''
}

geom_src = r'''
// ---- physical gradient from the transformed one ----------------------------
__device__ __forceinline__ void
transform_grad(fpdtype_t g[NDIMS][NVARS], const fpdtype_t s[NDIMS][NDIMS],
               const fpdtype_t rcpdjac)
{
    UNROLL for (int j = 0; j < NVARS; j++)
    {
        fpdtype_t t[NDIMS];
        UNROLL for (int k = 0; k < NDIMS; k++)
            t[k] = g[k][j];

        UNROLL for (int i = 0; i < NDIMS; i++)
        {
            fpdtype_t acc = 0;
            UNROLL for (int k = 0; k < NDIMS; k++)
                acc += s[k][i]*t[k];

            g[i][j] = rcpdjac*acc;
        }
    }
}

// ---- transformed flux: f[i][j] = sum_k s[i][k]*ft[k][j] ---------------------
__device__ __forceinline__ void
transform_flux(const fpdtype_t ft[NDIMS][NVARS],
               const fpdtype_t s[NDIMS][NDIMS], fpdtype_t f[NDIMS][NVARS])
{
    UNROLL for (int i = 0; i < NDIMS; i++)
        UNROLL for (int j = 0; j < NVARS; j++)
        {
            fpdtype_t acc = 0;
            UNROLL for (int k = 0; k < NDIMS; k++)
                acc += s[i][k]*ft[k][j];

            f[i][j] = acc;
        }
}
'''


def linear_smats_src(ndims, nverts, jac_exprs):
    """Metric terms of a linear element from its vertices, using the C
    Jacobian expressions the shape class supplies in ``tplargs``."""
    jl = '\n'.join(f'    j[{a}][{b}] = {jac_exprs[a][b]};'
                   for a in range(ndims) for b in range(ndims))

    if ndims == 2:
        body = '''
    s[0][0] =  j[1][1]; s[0][1] = -j[1][0];
    s[1][0] = -j[0][1]; s[1][1] =  j[0][0];
    d = s[0][0]*s[1][1] - s[0][1]*s[1][0];'''
    else:
        body = ''.join(f'''
    s[{i}][0] = j[{a}][1]*j[{b}][2] - j[{a}][2]*j[{b}][1];
    s[{i}][1] = j[{a}][2]*j[{b}][0] - j[{a}][0]*j[{b}][2];
    s[{i}][2] = j[{a}][0]*j[{b}][1] - j[{a}][1]*j[{b}][0];'''
                       for i, (a, b) in enumerate([(1, 2), (2, 0), (0, 1)]))
        body += '''
    d = j[0][0]*s[0][0] + j[0][1]*s[0][1] + j[0][2]*s[0][2];'''

    return f'''
__device__ __forceinline__ void
calc_smats_detj(const fpdtype_t V[{nverts}][NDIMS], const fpdtype_t x[NDIMS],
                fpdtype_t s[NDIMS][NDIMS], fpdtype_t &d)
{{
    fpdtype_t j[NDIMS][NDIMS];
{jl}
{body}
}}
'''


def smats_from_jac_src(ndims):
    """S-matrices (adjugate) and determinant from a Jacobian ``j[d][i] =
    d x_i/d xi_d`` -- the second half of ``calc_smats_detj``."""
    if ndims == 2:
        body = '''
    s[0][0] =  j[1][1]; s[0][1] = -j[1][0];
    s[1][0] = -j[0][1]; s[1][1] =  j[0][0];
    d = s[0][0]*s[1][1] - s[0][1]*s[1][0];'''
    else:
        body = ''.join(f'''
    s[{i}][0] = j[{a}][1]*j[{b}][2] - j[{a}][2]*j[{b}][1];
    s[{i}][1] = j[{a}][2]*j[{b}][0] - j[{a}][0]*j[{b}][2];
    s[{i}][2] = j[{a}][0]*j[{b}][1] - j[{a}][1]*j[{b}][0];'''
                       for i, (a, b) in enumerate([(1, 2), (2, 0), (0, 1)]))
        body += '''
    d = j[0][0]*s[0][0] + j[0][1]*s[0][1] + j[0][2]*s[0][2];'''

    return f'''
__device__ __forceinline__ void
smats_detj_from_jac(const fpdtype_t j[NDIMS][NDIMS],
                    fpdtype_t s[NDIMS][NDIMS], fpdtype_t &d)
{{
{body}
}}
'''


def multilinear_jacobian(jac_exprs, ndims, nverts):
    """Monomial form of the Jacobian expressions of a linear element.

    ``jac_exprs[d][i]`` (C expressions in the vertices ``V[n][i]`` and the
    reference point ``x[e]``, supplied by the shape class exactly as the
    reference's ``jac_exprs`` template argument) is linear in ``V`` with
    coefficients multilinear in ``x``.  Returns ``(monos, W)`` with
    ``monos`` a list of variable-index tuples (``()`` is the constant) and
    ``W[d][k][n]`` such that

        j[d][i] = sum_k prod(x[e] for e in monos[k]) * sum_n W[d][k][n]*V[n][i]

    or None when the expressions are not of that form.  The inner sums
    depend on the element only, so a kernel forms them once per element
    and evaluates the Jacobian at a point with a handful of FMAs instead
    of ``nverts*ndims`` products per entry."""
    import itertools as it

    import numpy as np

    monos = [m for r in range(ndims + 1)
             for m in it.combinations(range(ndims), r)]
    rng = np.random.default_rng(7)
    xs = rng.uniform(-1, 1, size=(4*len(monos), ndims))
    A = np.array([[np.prod(x[list(m)]) for m in monos] for x in xs])

    W = np.zeros((ndims, len(monos), nverts))
    for d in range(ndims):
        for n in range(nverts):
            cref = None
            for i in range(ndims):
                V = np.zeros((nverts, ndims))
                V[n][i] = 1.0
                try:
                    vals = np.array([eval(jac_exprs[d][i],
                                          {'V': V, 'x': x}) for x in xs])
                except Exception:
                    return None

                c, *_ = np.linalg.lstsq(A, vals, rcond=None)
                if np.abs(A @ c - vals).max() > 1e-12:
                    return None
                if cref is None:
                    cref = c
                elif np.abs(c - cref).max() > 1e-12:
                    return None

            cref[np.abs(cref) < 1e-12] = 0.0
            W[d, :, n] = cref

    # Coefficients of these maps are small dyadic rationals: snap them
    Wr = np.round(W*1024)/1024
    if np.abs(Wr - W).max() > 1e-12:
        return None

    return monos, Wr


def physics_defines(c, visc_corr='none', viscous=False):
    """``#define`` list for the gas constants a kernel needs."""
    d = [('C_GAMMA', fpconst(c['gamma']))]

    if viscous:
        d += [('C_MU', fpconst(c['mu'])),
              ('C_GAMMA_PR', fpconst(c['gamma']/c['Pr']))]

        if visc_corr == 'sutherland':
            d += [('VISC_SUTHERLAND', '1'),
                  ('C_RCPCPTREF', fpconst(1/c['cpTref'])),
                  ('C_MU_SUTH', fpconst(c['mu']*(c['cpTref'] + c['cpTs']))),
                  ('C_CPTS', fpconst(c['cpTs']))]

    return d
