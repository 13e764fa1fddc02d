"""The benchmark / parity workloads named in BASELINE.json, as config text
plus a synthetic mesh (PyFR-Test-Cases is not available offline).

``tgv``: 3-D Taylor-Green vortex, compressible Navier-Stokes, periodic hexes
(configs[1] and [2]); ``vortex``: 2-D Euler isentropic vortex, periodic
quads (configs[0]).  Settings follow SURVEY.md section 8(d).
"""

import numpy as np

from pyfr_b200.host.config import Config
from pyfr_b200.host.mesh import BoxMesh


def tgv_cfg(order=4, precision='double', rsolver='rusanov', beta=0.5,
            extra=''):
    return f'''
[backend]
precision = {precision}

[constants]
gamma = 1.4
mu = 6.25e-4
Pr = 0.71
M = 0.1

[solver]
system = navier-stokes
order = {order}
shock-capturing = none
viscosity-correction = none

[solver-interfaces]
riemann-solver = {rsolver}
ldg-beta = {beta}
ldg-tau = 0.1

[solver-interfaces-quad]
flux-pts = gauss-legendre

[solver-elements-hex]
soln-pts = gauss-legendre

[soln-ics]
rho = 1
u = sin(x)*cos(y)*cos(z)
v = -cos(x)*sin(y)*cos(z)
w = 0
p = 1/(gamma*M*M) + (cos(2*x) + cos(2*y))*(cos(2*z) + 2)/16
{extra}
'''


def vortex_cfg(order=3, precision='double', rsolver='rusanov', extra=''):
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
shock-capturing = none

[solver-interfaces]
riemann-solver = {rsolver}

[solver-interfaces-line]
flux-pts = gauss-legendre

[solver-elements-quad]
soln-pts = gauss-legendre

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


def make(case, n, **kw):
    if case == 'tgv':
        warp, curved = kw.pop('warp', 0.0), kw.pop('curved', 0.0)
        return Config(tgv_cfg(**kw)), tgv_mesh(n, warp, curved)
    elif case == 'vortex':
        return Config(vortex_cfg(**kw)), vortex_mesh(n)
    else:
        raise ValueError(f'Unknown case {case!r}')
