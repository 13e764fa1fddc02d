"""Host-side time integration (pyfr_b200/host/integrator.py) on the oracle
backend: formal order of the RK schemes and the PI controller's step-size
law (pyfr/integrators/explicit/steppers.py, controllers.py)."""

import math

import numpy as np
import pytest

from pyfr_b200 import cases
from pyfr_b200.host.integrator import PIController, RK4Stepper, RK45Stepper
from pyfr_b200.host.system import get_system

from util import OracleBackend, rel_err


def _run(stepper, dt, tend, **kw):
    cfg, box = cases.make('vortex', (3, 3), order=2)
    sysm = get_system(OracleBackend(cfg), box.local_mesh(), cfg, 4)
    st = stepper(sysm, **kw)
    st.advance(round(tend/dt), dt)
    return st.soln[0]


@pytest.mark.parametrize('stepper,kw', [(RK4Stepper, {}), (RK45Stepper, {}),
                                        (RK45Stepper, {'errest': True})],
                         ids=['rk4', 'rk45', 'rk45-errest'])
def test_fourth_order_in_time(stepper, kw):
    ref = _run(RK45Stepper, 0.0025, 0.16)
    e1 = rel_err(_run(stepper, 0.04, 0.16, **kw), ref)
    e2 = rel_err(_run(stepper, 0.02, 0.16, **kw), ref)

    assert 3.7 < math.log2(e1/e2) < 4.6, (e1, e2)


def test_rk45_coefficients_satisfy_order_conditions():
    """Butcher tableau of the 2R scheme: A[i][j] = b_j for j < i - 1,
    A[i][i-1] = a_{i-1}; main weights to order 4, embedded to order 3."""
    a, b, bh = RK45Stepper.a, RK45Stepper.b, RK45Stepper.bhat
    s = len(b)
    A = np.zeros((s, s))
    for i in range(1, s):
        A[i, :i - 1] = b[:i - 1]
        A[i, i - 1] = a[i - 1]
    c = A.sum(axis=1)

    for w, order in ((np.array(b), 4), (np.array(bh), 3)):
        conds = [(w.sum(), 1), (w @ c, 1/2), (w @ c**2, 1/3),
                 (w @ A @ c, 1/6)]
        if order == 4:
            conds += [(w @ c**3, 1/4), (w @ (c*(A @ c)), 1/8),
                      (w @ A @ c**2, 1/12), (w @ A @ A @ c, 1/24)]
        for got, want in conds:
            assert got == pytest.approx(want, abs=1e-12)

    assert abs(np.array(bh) @ c**3 - 1/4) > 1e-4     # genuinely third order


def test_pi_controller_step_size_law():
    cfg, box = cases.make('vortex', (3, 3), order=2)
    sysm = get_system(OracleBackend(cfg), box.local_mesh(), cfg, 4)
    sect = 'solver-time-integrator'
    for k, v in (('dt', 0.08), ('atol', 1e-6), ('rtol', 1e-6),
                 ('atol-rho', 1e-6)):
        cfg.set(sect, k, v)

    st = RK45Stepper(sysm, errest=True)
    with pytest.raises(ValueError, match='Missing atol'):
        PIController(st, cfg, ['rho', 'rhou', 'rhov', 'E'])

    for v in ('rhou', 'rhov', 'E'):
        cfg.set(sect, f'atol-{v}', 1e-6)
    pi = PIController(st, cfg, ['rho', 'rhou', 'rhov', 'E'])
    assert pi.gndofs == 9*9*4

    u0 = st.soln[0].copy()
    pi._dt_lookahead = 2
    pi.advance_to(0.3)
    assert pi.tcurr == 0.3 and pi.nrjctsteps > 0

    # dt_{n+1} = dt_n * clip(0.8 err^(-0.58/4) errprev^(0.42/4), 0.9, 1.1),
    # errprev updated on acceptance only; rejected steps leave t alone
    errprev, t = 1.0, 0.0
    for (dt, what, err), nxt in zip(pi.stepinfo, pi.stepinfo[1:]):
        fac = min(1.1, max(0.9, 0.8*err**(-0.58/4)*errprev**(0.42/4)))
        assert (what == 'accept') == (err < 1.0)
        if what == 'accept':
            errprev, t = err, t + dt
        # the clamp towards the end time may shorten the step
        assert nxt[0] <= fac*dt*(1 + 1e-14)
        if 0.3 - t > 3*fac*dt:
            assert nxt[0] == pytest.approx(fac*dt, rel=1e-14)

    assert sum(d for d, w, _ in pi.stepinfo if w == 'accept') == \
        pytest.approx(0.3, rel=1e-14)

    # a rejected first step restarts from the untouched initial state
    assert pi.stepinfo[0][1] == 'reject'
    sysm2 = get_system(OracleBackend(cfg), box.local_mesh(), cfg, 4)
    assert np.array_equal(sysm2.ele_scal_upts(0)[0], u0)

    with pytest.raises(ValueError, match='past'):
        pi.advance_to(0.1)


# -- against the reference's own integrators (tests/golden/make_golden.py) ------
def _golden_intg(name):
    import os
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, 'golden'))
    try:
        import make_golden as mg
    except ImportError:
        mg = None
    finally:
        sys.path.pop(0)

    return np.load(os.path.join(here, 'golden', f'intg_{name}.npz')), mg


INTG = {
    'vortex_p3_rk45_pi_l2': (
        'vortex', (4, 4), dict(order=3),
        dict(scheme='rk45', controller='pi', dt=0.05, atol=1e-6, rtol=1e-6),
        [0.2, 0.33]),
    'tgv_p2_rk45_pi_uniform': (
        'tgv', (2, 2, 2), dict(order=2),
        {'scheme': 'rk45', 'controller': 'pi', 'dt': 0.02,
         'errest-norm': 'uniform', 'rtol': 1e-5, 'atol-rho': 1e-5,
         'atol-rhou': 2e-5, 'atol-rhov': 2e-5, 'atol-rhow': 2e-5,
         'atol-E': 1e-4, 'pi-alpha': 0.7, 'pi-beta': 0.4,
         'safety-fact': 0.9, 'max-fact': 1.5, 'min-fact': 0.5,
         'dt-lookahead': 3},
        [0.15]),
    'vortex_p3_rk45_none': (
        'vortex', (4, 4), dict(order=3),
        dict(scheme='rk45', controller='none', dt=0.01), [0.05]),
    'vortex_p3_rk4_none': (
        'vortex', (4, 4), dict(order=3),
        dict(scheme='rk4', controller='none', dt=0.01), [0.05]),
    'vortex_p3_rk4_cfl': (
        'vortex', (4, 4), dict(order=3),
        {'scheme': 'rk4', 'controller': 'cfl', 'dt': 0.01, 'cfl': 0.4,
         'cfl-nsteps': 2}, [0.12]),
    'tgv_p2_rk45_cfl_curved': (
        'tgv', (2, 2, 2), dict(order=2, warp=0.1),
        {'scheme': 'rk45', 'controller': 'cfl', 'dt': 0.01, 'cfl': 0.3,
         'dt-max': 0.02}, [0.1]),
}


def make_integrator(sysm, cfg, opts):
    from pyfr_b200.host.integrator import CFLController, NoneController

    for k, v in opts.items():
        cfg.set('solver-time-integrator', k, v)

    pi = opts['controller'] == 'pi'
    st = (RK4Stepper(sysm) if opts['scheme'] == 'rk4' else
          RK45Stepper(sysm, errest=pi))

    if pi:
        nd = sysm.ndims
        convars = ['rho', 'rhou', 'rhov', 'rhow'][:nd + 1] + ['E']
        return PIController(st, cfg, convars), st
    elif opts['controller'] == 'cfl':
        return CFLController(st, cfg), st
    else:
        return NoneController(st, cfg), st


@pytest.mark.parametrize('name', list(INTG))
def test_host_integrators_reproduce_the_reference(name):
    """Same mesh, same oracle backend: the reference's composed integrator
    class (fixture) and the host mirror must take the same decisions and
    arrive at the same solution."""
    g, mg = _golden_intg(name)
    if mg is not None and hasattr(mg, 'INTG_CASES'):
        assert mg.INTG_CASES[name] == INTG[name], 'fixture recipe changed'

    case, n, kw, opts, tlist = INTG[name]
    cfg, box = cases.make(case, n, **kw)
    sysm = get_system(OracleBackend(cfg), box.local_mesh(), cfg, 4,
                      needs_cfl=opts['controller'] == 'cfl')
    ctl, st = make_integrator(sysm, cfg, opts)

    assert rel_err(st.soln[0], g['u0']) < 1e-14

    for i, t in enumerate(tlist):
        ctl.advance_to(t)
        assert ctl.tcurr == float(g[f'tcurr_t{i}'])
        assert rel_err(st.soln[0], g[f'u_t{i}']) < 2e-13

    hist = g['hist']
    assert [w == 'accept' for _, w, _ in ctl.stepinfo] == \
        [bool(a) for a in hist[:, 1]]
    np.testing.assert_allclose([d for d, _, _ in ctl.stepinfo], hist[:, 0],
                               rtol=1e-10)
    if opts['controller'] == 'pi':
        np.testing.assert_allclose([e for _, _, e in ctl.stepinfo],
                                   hist[:, 2], rtol=1e-8)
        assert ctl.gndofs == g['counts'][3]
        assert ctl.dt == pytest.approx(float(g['dt_final']), rel=1e-10)

    assert (ctl.nacptsteps, ctl.nrjctsteps) == tuple(g['counts'][:2])
