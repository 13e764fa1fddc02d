"""TEST ORACLE -- NumPy restatement of the reference's kernel arithmetic.

This file is test infrastructure, not product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import
it.  Each function restates one Mako macro/kernel of the reference
point-for-point (same operation order where it matters for rounding),
operating on lists of NumPy arrays, one array per variable.

PARITY STATUS: *pinned on the reference's own kernel templates.*  The
reference ships no golden vectors for flux, Riemann, gradient or RHS values
(only operator matrices, see tests/test_oracle_golden.py) and ``mako`` is
not installable here, so ``oracle/minimako.py`` renders the reference's
``.mako`` kernel files with a minimal interpreter of the template language
(macro expansion by the reference's own ``makoutil``); the rendered
kernels -- intcflux / mpicflux / intconu / mpiconu, bcconu / bccflux for
every boundary type, tflux in all its variants, gradcoru, wavespeed,
negdivconf, rkvdh2, fieldeval -- are compiled as C and this restatement
must reproduce them to 2e-13 on random states
(tests/test_oracle_templates.py); and with those kernels completed by the
reference's OpenMP generator and driven by the reference's host code
(oracle/refkernels.py) the whole RHS of the reference is reproduced
(tests/test_reference_kernels.py).  Beside that it is checked by analytic
properties (tests/test_oracle_physics.py: the viscous flux against the
analytic Newtonian stress tensor and Fourier heat flux, the inviscid flux
against the Euler flux, exact upwinding of HLLC for supersonic states,
impermeable and adiabatic walls, far-field transparency; tests/
test_host_logic.py: free-stream preservation, Riemann-solver consistency
f(u,u,n) = F(u).n and symmetry, discrete conservation, design order of
accuracy of the Euler RHS, convergence of the viscous RHS to an exact
solution) and by the reference's own host code driving it
(tests/golden/make_golden.py).
"""

import numpy as np


def _lsum(first, *rest):
    # Left-associated sum, the order a C compiler sees
    for r in rest:
        first = first + r
    return first


def inviscid_flux(s, ndims, nvars, c):
    """pyfr/solvers/euler/kernels/flux.mako:3-26 -> (f[d][v], p, v[d])"""
    invrho, E = 1.0/s[0], s[nvars - 1]

    rhov = [s[i + 1] for i in range(ndims)]
    v = [invrho*rv for rv in rhov]

    p = (c['gamma'] - 1)*(E - 0.5*invrho*sum(rv*rv for rv in rhov))

    f = [[None]*nvars for _ in range(ndims)]
    for i in range(ndims):
        f[i][0] = rhov[i]
        f[i][nvars - 1] = (E + p)*v[i]

    for i in range(ndims):
        for j in range(ndims):
            f[i][j + 1] = rhov[i]*v[j] + (p if i == j else 0.0)

    return f, p, v


def viscous_flux_add(u, gu, f, ndims, nvars, c, visc_corr='none'):
    """pyfr/solvers/navstokes/kernels/flux.mako:3-105; adds into f[d][v]"""
    gamma, mu, Pr = c['gamma'], c['mu'], c['Pr']
    rho, E = u[0], u[nvars - 1]
    rcprho = 1.0/rho
    vel = [rcprho*u[i + 1] for i in range(ndims)]

    rho_x = [gu[d][0] for d in range(ndims)]
    # dv[i][d] = rho * d(v_i)/d(x_d)
    dv = [[gu[d][i + 1] - vel[i]*rho_x[d] for d in range(ndims)]
          for i in range(ndims)]
    E_x = [gu[d][nvars - 1] for d in range(ndims)]

    if visc_corr == 'sutherland':
        cpT = gamma*(rcprho*E - 0.5*sum(v*v for v in vel))
        Trat = (1/c['cpTref'])*cpT
        mu_c = (mu*(c['cpTref'] + c['cpTs']))*Trat*np.sqrt(Trat) / \
               (cpT + c['cpTs'])
    else:
        mu_c = mu

    T_x = [rcprho*(E_x[d] - _lsum(rcprho*rho_x[d]*E,
                                  *(vel[i]*dv[i][d] for i in range(ndims))))
           for d in range(ndims)]

    div = sum(dv[i][i] for i in range(ndims))
    t = [[None]*ndims for _ in range(ndims)]
    for i in range(ndims):
        t[i][i] = -2*mu_c*rcprho*(dv[i][i] - (1.0/3.0)*div)
    for i in range(ndims):
        for j in range(i + 1, ndims):
            t[i][j] = t[j][i] = -mu_c*rcprho*(dv[j][i] + dv[i][j])

    for d in range(ndims):
        for i in range(ndims):
            f[d][i + 1] = f[d][i + 1] + t[d][i]

        f[d][nvars - 1] = f[d][nvars - 1] + _lsum(
            *(vel[i]*t[d][i] for i in range(ndims)),
            -mu_c*(gamma/Pr)*T_x[d]
        )


def rsolve_rusanov(ul, ur, n, ndims, nvars, c):
    """pyfr/solvers/euler/kernels/rsolvers/rusanov.mako:4-26"""
    fl, pl, vl = inviscid_flux(ul, ndims, nvars, c)
    fr, pr, vr = inviscid_flux(ur, ndims, nvars, c)

    nv = sum(n[i]*(vl[i] + vr[i]) for i in range(ndims))
    a = (np.sqrt((0.25*c['gamma'])*(pl + pr)/(ul[0] + ur[0]))
         + 0.25*np.abs(nv))

    return [0.5*sum(n[j]*(fl[j][i] + fr[j][i]) for j in range(ndims))
            + a*(ul[i] - ur[i]) for i in range(nvars)]


def rsolve_hllc(ul, ur, n, ndims, nvars, c):
    """pyfr/solvers/euler/kernels/rsolvers/hllc.mako:4-77"""
    gamma = c['gamma']
    fl, pl, vl = inviscid_flux(ul, ndims, nvars, c)
    fr, pr, vr = inviscid_flux(ur, ndims, nvars, c)

    nvl = sum(n[i]*vl[i] for i in range(ndims))
    nvr = sum(n[i]*vr[i] for i in range(ndims))

    al, ar = np.sqrt(gamma*pl/ul[0]), np.sqrt(gamma*pr/ur[0])
    srl, srr = np.sqrt(ul[0]), np.sqrt(ur[0])

    nv = (srl*nvl + srr*nvr)/(srl + srr)
    H = ((srl*(pr + ur[ndims + 1]) + srr*(pl + ul[ndims + 1]))
         / (srl*ur[0] + srr*ul[0]))

    inv_rar = 1/(srl + srr)
    va = [(vl[i]*srl + vr[i]*srr)*inv_rar for i in range(ndims)]
    qq = sum(v*v for v in va)
    a = np.sqrt((gamma - 1)*(H - 0.5*qq))

    sl = np.minimum(nv - a, nvl - al)
    sr = np.maximum(nv + a, nvr + ar)
    sstar = ((pr - pl + ul[0]*nvl*(sl - nvl) - ur[0]*nvr*(sr - nvr)) /
             (ul[0]*(sl - nvl) - ur[0]*(sr - nvr)))

    ul_com = (sl - nvl)/(sl - sstar)
    ur_com = (sr - nvr)/(sr - sstar)

    usl, usr = [None]*nvars, [None]*nvars
    usl[0], usr[0] = ul_com*ul[0], ur_com*ur[0]
    for i in range(ndims):
        usl[i + 1] = usl[0]*(vl[i] + (sstar - nvl)*n[i])
        usr[i + 1] = usr[0]*(vr[i] + (sstar - nvr)*n[i])

    usl[nvars - 1] = ul_com*(ul[nvars - 1] + (sstar - nvl) *
                             (ul[0]*sstar + pl/(sl - nvl)))
    usr[nvars - 1] = ur_com*(ur[nvars - 1] + (sstar - nvr) *
                             (ur[0]*sstar + pr/(sr - nvr)))

    nf = []
    for i in range(nvars):
        nf_fl = sum(n[j]*fl[j][i] for j in range(ndims))
        nf_fr = sum(n[j]*fr[j][i] for j in range(ndims))
        nf_fsl = nf_fl + sl*(usl[i] - ul[i])
        nf_fsr = nf_fr + sr*(usr[i] - ur[i])

        nf.append(np.where(
            0 <= sl, nf_fl, np.where(
                (sl <= 0) & (0 <= sstar), nf_fsl, np.where(
                    (sstar <= 0) & (0 <= sr), nf_fsr, nf_fr))))

    return nf


rsolvers = {'rusanov': rsolve_rusanov, 'hllc': rsolve_hllc}


def calc_smats_detj(jac_exprs, V, x, ndims):
    """pyfr/solvers/baseadvec/kernels/smats.mako:3-25; the Jacobian
    expressions are the C strings of the shape class, which are also valid
    Python over arrays."""
    env = {'__builtins__': {}, 'x': x, 'V': V}
    j = [[eval(jac_exprs[a][b], env) for b in range(ndims)]
         for a in range(ndims)]

    if ndims == 2:
        s = [[j[1][1], -j[1][0]], [-j[0][1], j[0][0]]]
        d = s[0][0]*s[1][1] - s[0][1]*s[1][0]
    else:
        s = []
        for a, b in [(1, 2), (2, 0), (0, 1)]:
            s.append([j[a][1]*j[b][2] - j[a][2]*j[b][1],
                      j[a][2]*j[b][0] - j[a][0]*j[b][2],
                      j[a][0]*j[b][1] - j[a][1]*j[b][0]])
        d = j[0][0]*s[0][0] + j[0][1]*s[0][1] + j[0][2]*s[0][2]

    return s, d


def transform_grad(g, smats, rcpdjac, ndims, nvars):
    """pyfr/solvers/baseadvecdiff/kernels/transform_grad.mako:3-10"""
    return [[rcpdjac*sum(smats[k][i]*g[k][j] for k in range(ndims))
             for j in range(nvars)] for i in range(ndims)]


def transform_flux(ft, smats, ndims, nvars):
    """Last loop of tflux.mako: f[i][j] = sum_k smats[i][k]*ftemp[k][j]"""
    return [[sum(smats[i][k]*ft[k][j] for k in range(ndims))
             for j in range(nvars)] for i in range(ndims)]


def ns_common_flux(ul, ur, gul, gur, nl, ndims, nvars, c, rsolver,
                   visc_corr='none'):
    """pyfr/solvers/navstokes/kernels/intcflux.mako:10-54 (and mpicflux):
    returns the un-signed common normal flux mag_nl*(ficomm + fvcomm)."""
    beta, tau = c['ldg-beta'], c['ldg-tau']

    mag_nl = np.sqrt(sum(x*x for x in nl))
    n = [(1/mag_nl)*x for x in nl]

    ficomm = rsolvers[rsolver](ul, ur, n, ndims, nvars, c)

    zeros = lambda: [[0.0]*nvars for _ in range(ndims)]
    if beta != -0.5:
        fvl = zeros()
        viscous_flux_add(ul, gul, fvl, ndims, nvars, c, visc_corr)
    if beta != 0.5:
        fvr = zeros()
        viscous_flux_add(ur, gur, fvr, ndims, nvars, c, visc_corr)

    out = []
    for i in range(nvars):
        if beta == -0.5:
            fv = sum(n[j]*fvr[j][i] for j in range(ndims))
        elif beta == 0.5:
            fv = sum(n[j]*fvl[j][i] for j in range(ndims))
        else:
            fv = ((0.5 + beta)*sum(n[j]*fvl[j][i] for j in range(ndims))
                  + (0.5 - beta)*sum(n[j]*fvr[j][i] for j in range(ndims)))

        if tau != 0.0:
            fv = fv + tau*(ul[i] - ur[i])

        out.append(mag_nl*(ficomm[i] + fv))

    return out


def euler_common_flux(ul, ur, nl, ndims, nvars, c, rsolver):
    """pyfr/solvers/euler/kernels/intcflux.mako:8-24"""
    mag_nl = np.sqrt(sum(x*x for x in nl))
    n = [(1/mag_nl)*x for x in nl]
    fn = rsolvers[rsolver](ul, ur, n, ndims, nvars, c)

    return [mag_nl*f for f in fn]


def ldg_common_solution(ul, ur, beta):
    """navstokes/kernels/intconu.mako:4-19 -> (ulout, urout); None means the
    side is not written."""
    if beta == -0.5:
        return None, ul
    elif beta == 0.5:
        return ur, None
    else:
        com = [r*(0.5 + beta) + l*(0.5 - beta) for l, r in zip(ul, ur)]
        return com, com


# -- boundary conditions --------------------------------------------------------
def _bcval(c, key, env):
    """A BC constant: a number, or a C expression string (already
    substituted by the host, pyfr/solvers/baseadvec/inters.py:98-119) in
    terms of ``ploc[i]`` and ``t``."""
    v = c[key]
    if isinstance(v, str):
        fns = {'sqrt': np.sqrt, 'exp': np.exp, 'log': np.log, 'sin': np.sin,
               'cos': np.cos, 'tan': np.tan, 'tanh': np.tanh, 'pow': np.power,
               'fabs': np.abs}
        return eval(v, {'__builtins__': {}}, dict(fns, **env))
    return v


def _ke2(u, ndims):
    return sum(u[i + 1]*u[i + 1] for i in range(ndims))


def bc_rsolve_state(bctype, ul, nl, ndims, nvars, c, env):
    """``bc_rsolve_state`` of pyfr/solvers/{euler,navstokes}/kernels/bcs/
    <bctype>.mako: the ghost state handed to the Riemann solver."""
    gamma = c['gamma']
    gmo = gamma - 1.0
    uvw = 'uvw'[:ndims]

    if bctype == 'no-slp-adia-wall':
        return [ul[0], *(-ul[i + 1] for i in range(ndims)), ul[nvars - 1]]
    elif bctype == 'slp-adia-wall':
        nor = sum(ul[i + 1]*nl[i] for i in range(ndims))
        return [ul[0], *(ul[i + 1] - 2*nor*nl[i] for i in range(ndims)),
                ul[nvars - 1]]
    elif bctype == 'no-slp-isot-wall':
        ur = [ul[0]] + [-ul[i + 1] + 2*_bcval(c, v, env)*ul[0]
                        for i, v in enumerate(uvw)]
        ur.append((c['cpTw']/gamma)*ur[0] + 0.5*(1.0/ur[0])*_ke2(ur, ndims))
        return ur
    elif bctype == 'sup-out-fn':
        return list(ul)
    elif bctype == 'sup-in-fa':
        rho = _bcval(c, 'rho', env)
        ur = [rho + 0*ul[0]] + [rho*_bcval(c, v, env) + 0*ul[0] for v in uvw]
        ur.append(_bcval(c, 'p', env)/gmo + 0.5*(1.0/ur[0])*_ke2(ur, ndims))
        return ur
    elif bctype == 'sub-in-frv':
        rho = _bcval(c, 'rho', env)
        ur = [rho + 0*ul[0]] + [rho*_bcval(c, v, env) + 0*ul[0] for v in uvw]
        ur.append(ul[nvars - 1] - 0.5*(1.0/ul[0])*_ke2(ul, ndims)
                  + 0.5*(1.0/ur[0])*_ke2(ur, ndims))
        return ur
    elif bctype == 'sub-out-fp':
        return [*ul[:nvars - 1],
                _bcval(c, 'p', env)/gmo + 0.5*(1.0/ul[0])*_ke2(ul, ndims)]
    elif bctype == 'sub-in-ftpttang':
        # navstokes/kernels/bcs/sub-in-ftpttang.mako: total pressure and
        # temperature with a prescribed flow direction
        pl = gmo*(ul[nvars - 1] - (0.5/ul[0])*_ke2(ul, ndims))
        udotu = (2.0*c['cpTt'])*(1.0 - c['pt']**(-c['Rdcp'])*pl**c['Rdcp'])
        udotu = np.maximum(0, udotu)

        ur = [(1.0/c['Rdcp'])*pl/(c['cpTt'] - 0.5*udotu)]
        ur += [v*ur[0]*np.sqrt(udotu) for v in c['vc']]
        ur.append((1.0/gmo)*pl + 0.5*ur[0]*udotu)
        return ur
    elif bctype == 'char-riem-inv':
        pe, rhoe = _bcval(c, 'p', env), _bcval(c, 'rho', env)
        ve = [_bcval(c, v, env) for v in uvw]

        cs = np.sqrt(gamma*pe/rhoe)
        s = pe*rhoe**(-gamma)
        ratio = cs*(2.0/gmo)

        inv = 1.0/ul[0]
        V_e = sum(ve[i]*nl[i] for i in range(ndims))
        V_i = inv*sum(ul[i + 1]*nl[i] for i in range(ndims))
        p_i = gmo*ul[nvars - 1] - (0.5*gmo)*inv*_ke2(ul, ndims)
        c_i = np.sqrt(gamma*p_i*inv)

        sup = np.abs(V_e) >= cs
        R_e = np.where(sup & (V_i >= 0), V_i - c_i*(2.0/gmo), V_e - ratio)
        R_i = np.where(sup & (V_i < 0), V_e + ratio, V_i + c_i*(2.0/gmo))
        V_b = 0.5*(R_e + R_i)
        c_b = (0.25*gmo)*(R_i - R_e)
        rho_b = np.where(
            V_i < 0, ((1.0/(gamma*s))*c_b*c_b)**(1.0/gmo),
            ul[0]*(ul[0]*c_b*c_b/(gamma*p_i))**(1.0/gmo)
        )
        p_b = (1.0/gamma)*rho_b*c_b*c_b

        ur = [rho_b]
        for i in range(ndims):
            ur.append(np.where(
                V_i >= 0, rho_b*(ul[i + 1]*inv + (V_b - V_i)*nl[i]),
                rho_b*(ve[i] + (V_b - V_e)*nl[i])
            ))
        ur.append(p_b*(1.0/gmo) + 0.5*(1.0/ur[0])*_ke2(ur, ndims))
        return ur
    else:
        raise NotImplementedError(f'oracle: boundary type {bctype!r}')


def bc_ldg_state(bctype, ul, nl, ndims, nvars, c, env):
    """``bc_ldg_state`` (navstokes/kernels/bcs/<bctype>.mako): the state
    the LDG common solution / viscous ghost flux is built from."""
    if bctype == 'no-slp-adia-wall':
        return [ul[0], *(0.0*ul[0] for _ in range(ndims)),
                ul[nvars - 1] - (0.5/ul[0])*_ke2(ul, ndims)]
    elif bctype == 'no-slp-isot-wall':
        ur = [ul[0]] + [_bcval(c, v, env)*ul[0] for v in 'uvw'[:ndims]]
        ur.append((c['cpTw']/c['gamma'])*ur[0]
                  + 0.5*(1.0/ur[0])*_ke2(ur, ndims))
        return ur
    else:
        # aliased to bc_rsolve_state for every other type
        return bc_rsolve_state(bctype, ul, nl, ndims, nvars, c, env)


def bc_ldg_grad_state(bctype, ur, nl, gul, ndims, nvars):
    """``bc_ldg_grad_state``: gradient used on the ghost side."""
    if bctype in ('char-riem-inv', 'sup-in-fa', 'sub-in-frv', 'sub-out-fp'):
        return [[0.0*g for g in row] for row in gul]
    elif bctype in ('sup-out-fn', 'no-slp-isot-wall', 'sub-in-ftpttang'):
        return [list(row) for row in gul]
    elif bctype == 'no-slp-adia-wall':
        # no-slp-adia-wall.mako:22-75: remove the wall-normal temperature
        # gradient from the copied gradients
        rcprho = 1.0/ur[0]
        vel = [rcprho*ur[i + 1] for i in range(ndims)]
        dv = [[gul[d][i + 1] - vel[i]*gul[d][0] for d in range(ndims)]
              for i in range(ndims)]
        Tl = [gul[d][nvars - 1] - _lsum(rcprho*gul[d][0]*ur[nvars - 1],
                                        *(vel[i]*dv[i][d]
                                          for i in range(ndims)))
              for d in range(ndims)]

        gur = [list(row) for row in gul]
        for d in range(ndims):
            gur[d][nvars - 1] = gur[d][nvars - 1] - _lsum(
                *(nl[d]*nl[e]*Tl[e] for e in range(ndims))
            )
        return gur
    else:
        raise NotImplementedError(f'oracle: boundary type {bctype!r}')


def bc_common_flux(bctype, cflux_state, ul, gul, nl, ndims, nvars, c,
                   rsolver, env, viscous, visc_corr='none'):
    """bccflux kernels: euler/kernels/bccflux.mako and, for the viscous
    system, ``bc_common_flux_state`` of navstokes/kernels/bcs/{ghost,
    ghost-imperm}.mako or of the type's own template (slp-adia-wall)."""
    mag = np.sqrt(sum(x*x for x in nl))
    n = [(1/mag)*x for x in nl]

    if not viscous or cflux_state is None:
        ur = bc_rsolve_state(bctype, ul, n, ndims, nvars, c, env)
        fn = rsolvers[rsolver](ul, ur, n, ndims, nvars, c)
        return [mag*f for f in fn]

    ur = bc_ldg_state(bctype, ul, n, ndims, nvars, c, env)
    # ghost.mako passes ul, ghost-imperm.mako ur, as the state argument
    gur = bc_ldg_grad_state(bctype, ul if cflux_state == 'ghost' else ur,
                            n, gul, ndims, nvars)

    fvr = [[0.0]*nvars for _ in range(ndims)]
    viscous_flux_add(ur, gur, fvr, ndims, nvars, c, visc_corr)

    ur = bc_rsolve_state(bctype, ul, n, ndims, nvars, c, env)
    fi = rsolvers[rsolver](ul, ur, n, ndims, nvars, c)

    out = []
    for i in range(nvars):
        fv = sum(n[j]*fvr[j][i] for j in range(ndims))
        if cflux_state == 'ghost' and c['ldg-tau'] != 0.0:
            fv = fv + c['ldg-tau']*(ul[i] - ur[i])
        out.append(mag*(fi[i] + fv))

    return out
