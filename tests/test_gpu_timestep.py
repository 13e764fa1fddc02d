"""GPU parity of the callers either side of the RHS: RK4 steps (4 RHS +
6 axnpby per step) and the ``integrate`` diagnostics (Taylor-Green kinetic
energy and enstrophy via ``compute_grads`` + ``fieldeval``), B200 backend
against the NumPy oracle driven by the same host code.

BASELINE.json's acceptance: integrated kinetic energy and enstrophy within
1e-9 relative after 1000 steps (fp64)."""

import numpy as np
import pytest

from pyfr_b200 import cases
from pyfr_b200.host.integrator import (FieldIntegrator, RK4Stepper,
                                       TGV_EXPRS)
from pyfr_b200.host.system import get_system

from util import OracleBackend, rel_err

pytestmark = pytest.mark.gpu


def _run(which, n, nsteps, dt, sample=(), **kw):
    from pyfr_b200.backend import B200Backend

    cfg, box = cases.make('tgv', n, **kw)
    if which == 'b200':
        be = B200Backend(cfg)
    elif which == 'c':
        import os

        from oracle.cbackend import make_cbackend
        from pyfr_b200 import base
        be = make_cbackend(base, fast=False, nthreads=os.cpu_count())(cfg)
    else:
        be = OracleBackend(cfg)
    sysm = get_system(be, box.local_mesh(), cfg, 3)
    fi = FieldIntegrator(sysm, cfg, TGV_EXPRS)
    st = RK4Stepper(sysm)

    hist = [fi(st.tcurr, st.idxcurr)]
    for i in range(1, nsteps + 1):
        st.step(dt)
        if i in sample or i == nsteps:
            hist.append(fi(st.tcurr, st.idxcurr))

    return np.array(hist), st.soln[0]


def test_integrals_at_t0_match_analytic(built):
    """K = 1/8 (2 pi)^3 rho0 and enstrophy = 3/8 (2 pi)^3 for the TGV
    initial condition (quadrature of a p=4 interpolant on 6^3 elements)."""
    hist, _ = _run('b200', 6, 0, 0.0, order=4)
    vol = (2*np.pi)**3

    assert abs(hist[0][0]/vol - 0.125) < 1e-6
    assert abs(hist[0][1]/vol - 0.375) < 5e-4


@pytest.mark.parametrize('kw', [dict(order=3, warp=0.1),
                                dict(order=2, rsolver='hllc', beta=0.0)],
                         ids=str)
def test_rk4_steps_match_oracle(built, kw):
    n, nsteps, dt = (4, 3, 3), 25, 2e-3
    ho, so = _run('oracle', n, nsteps, dt, sample=(1, 10), **kw)
    hb, sb = _run('b200', n, nsteps, dt, sample=(1, 10), **kw)

    assert rel_err(sb, so) < 1e-11
    assert np.abs(hb/ho - 1).max() < 1e-11


def test_tgv_integrals_after_1000_steps(built):
    n, nsteps, dt = 4, 1000, 2e-3
    ho, so = _run('oracle', n, nsteps, dt, sample=(250, 500), order=2)
    hb, sb = _run('b200', n, nsteps, dt, sample=(250, 500), order=2)

    # kinetic energy and enstrophy histories
    assert np.abs(hb/ho - 1).max() < 1e-9, (hb, ho)
    # the flow has evolved (this is not a comparison of initial states)
    assert ho[-1][0] < 0.999*ho[0][0]
    assert rel_err(sb, so) < 1e-9


def test_tgv_integrals_after_1000_steps_p4(built):
    """BASELINE.json's acceptance run at the headline order: 16^3 hexes,
    p = 4, fp64, 1000 RK4 steps (4000 RHS evaluations of 2.56 M DoF)
    against the C restatement of the path (oracle/crhs, no -ffast-math):
    integrated kinetic energy and enstrophy within 1e-9 relative."""
    n, nsteps, dt = 16, 1000, 1e-3
    ho, so = _run('c', n, nsteps, dt, sample=(250, 500), order=4)
    hb, sb = _run('b200', n, nsteps, dt, sample=(250, 500), order=4)

    dev = float(np.abs(hb/ho - 1).max())
    from util import PARITY_LOG
    PARITY_LOG.append(dict(test='TGV 16^3 p=4, 1000 RK4 steps: KE/enstrophy '
                                'vs oracle/crhs', err=dev, floor=0.0,
                           ratio=None, ratio_oracle=None,
                           soln_err=float(rel_err(sb, so))))

    assert dev < 1e-9, (hb, ho)
    assert ho[-1][0] < 0.999*ho[0][0]
    assert rel_err(sb, so) < 1e-9
