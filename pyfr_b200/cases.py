"""The benchmark / parity workloads named in BASELINE.json, as config text
plus a synthetic mesh (PyFR-Test-Cases is not available offline).

``tgv``: 3-D Taylor-Green vortex, compressible Navier-Stokes, periodic hexes
(configs[1] and [2]); ``vortex``: 2-D Euler isentropic vortex, periodic
quads (configs[0]).  Settings follow SURVEY.md section 8(d).
"""

import numpy as np

from pyfr_b200.host.config import Config
from pyfr_b200.host.mesh import BoxMesh, MixedBoxMesh


def tgv_cfg(order=4, precision='double', rsolver='rusanov', beta=0.5,
            extra='', antialias='none', visc_corr='none',
            pts='gauss-legendre'):
    return f'''
[backend]
precision = {precision}

[constants]
gamma = 1.4
mu = 6.25e-4
Pr = 0.71
M = 0.1
cpTref = 250.0
cpTs = 95.0

[solver]
system = navier-stokes
order = {order}
anti-alias = {antialias}
shock-capturing = none
viscosity-correction = {visc_corr}

[solver-interfaces]
riemann-solver = {rsolver}
ldg-beta = {beta}
ldg-tau = 0.1

[solver-interfaces-quad]
flux-pts = {pts}

[solver-elements-hex]
soln-pts = {pts}

[soln-ics]
rho = 1
u = sin(x)*cos(y)*cos(z)
v = -cos(x)*sin(y)*cos(z)
w = 0
p = 1/(gamma*M*M) + (cos(2*x) + cos(2*y))*(cos(2*z) + 2)/16
{extra}
'''


def vortex_cfg(order=3, precision='double', rsolver='rusanov', extra='',
               antialias='none', pts='gauss-legendre'):
    return f'''
[backend]
precision = {precision}

[constants]
gamma = 1.4
S = 13.5
M = 0.4
R = 1.5

[solver]
system = euler
order = {order}
anti-alias = {antialias}
shock-capturing = none

[solver-interfaces]
riemann-solver = {rsolver}

[solver-interfaces-line]
flux-pts = {pts}

[solver-elements-quad]
soln-pts = {pts}

[soln-ics]
rho = pow(1 - S*S*M*M*(gamma - 1)*exp(2*(1 - x*x - y*y)/(2*R*R))/(8*pi*pi), 1/(gamma - 1))
u = S*y*exp((1 - x*x - y*y)/(2*R*R))/(2*pi*R)
v = 1 - S*x*exp((1 - x*x - y*y)/(2*R*R))/(2*pi*R)
p = pow(1 - S*S*M*M*(gamma - 1)*exp(2*(1 - x*x - y*y)/(2*R*R))/(8*pi*pi), gamma/(gamma - 1))/(gamma*M*M)
{extra}
'''


def tgv_mesh(n, warp=0.0, curved=0.0):
    n = (n,)*3 if np.isscalar(n) else n
    return BoxMesh(n, -np.pi, np.pi, periodic=True, warp=warp, curved=curved)


def vortex_mesh(n=40):
    n = (n,)*2 if np.isscalar(n) else n
    return BoxMesh(n, -20.0, 20.0, periodic=True)


# Boundary sections used by the wall-bounded parity cases
BC_SECTIONS = {
    'no-slp-adia-wall': 'type = no-slp-adia-wall\n',
    'slp-adia-wall': 'type = slp-adia-wall\n',
    'no-slp-isot-wall': 'type = no-slp-isot-wall\ncpTw = 250.0\nu = 0.1\n',
    'char-riem-inv': ('type = char-riem-inv\nrho = 1.0\nu = 0.2\nv = 0.1\n'
                      'w = 0.0\np = 71.0\n'),
    'sup-out-fn': 'type = sup-out-fn\n',
    'sup-in-fa': ('type = sup-in-fa\nrho = 1.0\nu = 0.2\nv = 0.1\nw = 0.0\n'
                  'p = 71.0\n'),
    'sub-in-frv': 'type = sub-in-frv\nrho = 1.0\nu = 0.2\nv = 0.1\nw = 0.0\n',
    'sub-out-fp': 'type = sub-out-fp\np = 71.0\n',
    'sub-in-ftpttang': ('type = sub-in-ftpttang\npt = 75.0\ncpTt = 260.0\n'
                        'theta = 20.0\nphi = 80.0\n'),
}


def box_case(system, n, bcs, order=3, rsolver='rusanov', beta=0.5,
             precision='double', warp=0.0):
    """A box with boundaries: ``bcs`` maps boundary names (``'xlo'``,
    ``'yhi'`` ...) to boundary types; axes without an entry stay periodic.
    Navier-Stokes in 3-D (TGV initial condition), Euler in 2-D (vortex)."""
    nd = 3 if system == 'navier-stokes' else 2
    n = (n,)*nd if np.isscalar(n) else n
    periodic = tuple(not any(b.startswith('xyz'[a]) for b in bcs)
                     for a in range(nd))

    # A smooth field with a mean flow through every face: boundary types
    # that switch on the sign of the normal velocity (char-riem-inv) are
    # then evaluated away from their discontinuity
    if nd == 3:
        txt = tgv_cfg(order=order, precision=precision, rsolver=rsolver,
                      beta=beta)
        ics = ('rho = 1 + 0.1*sin(x)*cos(y)\nu = 0.3 + 0.1*cos(x + y)\n'
               'v = 0.15 + 0.1*sin(y)*cos(z)\nw = 0.1 + 0.05*sin(z + x)\n'
               'p = 71*(1 + 0.02*cos(x)*sin(z))\n')
        box = BoxMesh(n, -np.pi, np.pi, periodic=periodic, warp=warp)
    else:
        txt = vortex_cfg(order=order, precision=precision, rsolver=rsolver)
        ics = ('rho = 1 + 0.1*sin(0.3*x)*cos(0.2*y)\n'
               'u = 0.3 + 0.1*cos(0.2*(x + y))\nv = 0.15 + 0.1*sin(0.3*y)\n'
               'p = 4.5*(1 + 0.02*cos(0.2*x))\n')
        box = BoxMesh(n, -20.0, 20.0, periodic=periodic, warp=warp)

    txt = txt[:txt.index('[soln-ics]')] + '[soln-ics]\n' + ics

    for name, btype in bcs.items():
        sect = BC_SECTIONS[btype]
        if nd == 2:
            sect = sect.replace('w = 0.0\n', '')
        txt += f'\n[soln-bcs-{name}]\n{sect}'

    return Config(txt), box, txt


# Point sets of BASELINE.json configs[3]
MIXED_POINTS = '''
[solver-interfaces-line]
flux-pts = gauss-legendre
[solver-interfaces-quad]
flux-pts = gauss-legendre
[solver-interfaces-tri]
flux-pts = williams-shunn
[solver-elements-quad]
soln-pts = gauss-legendre
[solver-elements-tri]
soln-pts = williams-shunn
[solver-elements-hex]
soln-pts = gauss-legendre
[solver-elements-tet]
soln-pts = shunn-ham
[solver-elements-pri]
soln-pts = williams-shunn~gauss-legendre
[solver-elements-pyr]
soln-pts = gauss-legendre
'''

MIXED_PATTERNS = {
    'quad+tri': ['quad', 'tri'],
    'hex+pri': ['hex', 'pri'],
    'hex+pri+pyr+tet': ['hex', 'pri', 'pyr', 'pyt'],
}


def mixed_case(pattern, n, order=3, warp=0.05, h=1.0, **kw):
    """A periodic box mixing element types (``pattern``: a key of
    ``MIXED_PATTERNS``; ``n = (nx, ny[, nz])`` unit cells): Euler in 2-D,
    Navier-Stokes in 3-D, smooth box-periodic initial condition.  Returns
    ``(Config, MixedBoxMesh, config text)``."""
    import re

    kinds = MixedBoxMesh.columns(*(tuple(n) + (None,))[:3],
                                 MIXED_PATTERNS[pattern])
    nd = kinds.ndim

    if nd == 2:
        txt = vortex_cfg(order=order, **kw)
        ics = ('rho = 1 + 0.1*sin(kx*x)*cos(ky*y)\n'
               'u = 0.3 + 0.1*cos(kx*x + ky*y)\nv = 0.15 + 0.1*sin(ky*y)\n'
               'p = 4.5*(1 + 0.02*cos(kx*x))\n')
    else:
        txt = tgv_cfg(order=order, **kw)
        ics = ('rho = 1 + 0.1*sin(kx*x)*cos(ky*y)\n'
               'u = 0.3 + 0.1*cos(kx*x + ky*y)\n'
               'v = 0.15 + 0.1*sin(ky*y)*cos(kz*z)\n'
               'w = 0.1 + 0.05*sin(kz*z + kx*x)\n'
               'p = 71*(1 + 0.02*cos(kx*x)*sin(kz*z))\n')

    ks = ''.join(f'k{"xyz"[a]} = {2*np.pi/(kinds.shape[a]*h)!r}\n'
                 for a in range(nd))
    head = txt.partition('[soln-ics]')[0]
    head = head.replace('[constants]\n', '[constants]\n' + ks)

    extra = MIXED_POINTS
    for sect in re.findall(r'^\[([^\]]+)\]', head, flags=re.M):
        extra = re.sub(r'\[' + re.escape(sect) + r'\]\n[^\[]*', '', extra)

    txt = head + '[soln-ics]\n' + ics + extra
    return Config(txt), MixedBoxMesh(kinds, h=h, warp=warp), txt


def make(case, n, **kw):
    if case == 'tgv':
        warp, curved = kw.pop('warp', 0.0), kw.pop('curved', 0.0)
        return Config(tgv_cfg(**kw)), tgv_mesh(n, warp, curved)
    elif case == 'vortex':
        return Config(vortex_cfg(**kw)), vortex_mesh(n)
    else:
        raise ValueError(f'Unknown case {case!r}')
