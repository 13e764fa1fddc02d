"""CPU tests: the C11/OpenMP restatement (oracle/crhs, the CPU baseline
bench.py times) against the NumPy oracle on identical inputs -- two
independent restatements of the reference's kernel arithmetic must agree
to round-off, single- and multi-partition, with the block-group fusion and
thread-local scratch substitution of the reference's OpenMP design."""

import numpy as np
import pytest

from oracle.cbackend import make_cbackend
from oracle.npbackend import LocalComm
from pyfr_b200 import base, cases
from pyfr_b200.host.system import get_system

from util import oracle_rhs, rel_err, run_lockstep


@pytest.fixture(scope='module')
def CBackend():
    return make_cbackend(base)


def c_rhs(CBackend, case, n, parts=None, **kw):
    nparts = int(np.prod(parts)) if parts else 1
    world = LocalComm(0, nparts)
    systems = []

    for r in range(nparts):
        cfg, box = cases.make(case, n, **kw)
        vparts = box.brick_partition(parts) if nparts > 1 else None
        systems.append(get_system(CBackend(cfg), box.local_mesh(vparts, r),
                                  cfg, 2, comm=world.peer(r)))

    run_lockstep(systems, world, 0.0, 0, 1)
    return systems, [s.ele_scal_upts(1)[0] for s in systems]


@pytest.mark.parametrize('case,n,kw', [
    ('tgv', (3, 2, 2), dict(order=2, warp=0.1)),
    ('tgv', (3, 2, 2), dict(order=3, rsolver='hllc', beta=0.0, warp=0.1)),
    ('tgv', (2, 2, 3), dict(order=4, beta=-0.5)),
    ('vortex', 6, dict(order=3)),
    ('vortex', 5, dict(order=2, rsolver='hllc')),
])
def test_c_backend_matches_numpy_oracle(CBackend, case, n, kw):
    systems, out = c_rhs(CBackend, case, n, **kw)
    _, ref = oracle_rhs(case, n, **kw)

    assert rel_err(out[0], ref[0]) < 1e-12

    # The element chains really ran as fused block groups in C
    names = [type(k).__name__ for g in systems[0].rhs_graphs(0, 1)
             for w, k in g.program if w == 'kernel']
    assert 'GroupKernel' in names and 'NPKernel' not in names


@pytest.mark.parametrize('case,n,parts,kw', [
    ('tgv', (4, 2, 2), (2, 1, 1), dict(order=2, warp=0.1)),
    ('tgv', (4, 4, 2), (2, 2, 1), dict(order=1, beta=0.0, rsolver='hllc')),
    ('vortex', (6, 4), (2, 1), dict(order=3)),
])
def test_c_backend_partitioned(CBackend, case, n, parts, kw):
    _, out = c_rhs(CBackend, case, n, parts=parts, **kw)

    _, box = cases.make(case, n, **kw)
    vparts = box.brick_partition(parts)
    _, ref = oracle_rhs(case, n, vparts=vparts, nparts=len(out), **kw)

    for o, r in zip(out, ref):
        assert rel_err(o, r) < 1e-12


def test_rk4_and_integrals_c_vs_numpy(CBackend):
    """RK4 steps + Taylor-Green diagnostics: the two CPU restatements stay
    together over time (also covers axnpby, compute_grads, fieldeval)."""
    from pyfr_b200.host.integrator import (FieldIntegrator, RK4Stepper,
                                           TGV_EXPRS)
    from util import OracleBackend

    res = []
    for cls in (OracleBackend, CBackend):
        cfg, box = cases.make('tgv', (3, 3, 2), order=2, warp=0.1)
        s = get_system(cls(cfg), box.local_mesh(), cfg, 3)
        fi, st = FieldIntegrator(s, cfg, TGV_EXPRS), RK4Stepper(s)
        h = [fi(0.0, st.idxcurr)]
        st.advance(10, 2e-3)
        h.append(fi(st.tcurr, st.idxcurr))
        res.append((np.array(h), st.soln[0]))

    (h0, s0), (h1, s1) = res
    assert np.abs(h1/h0 - 1).max() < 1e-12
    assert rel_err(s1, s0) < 1e-12
    assert abs(h0[0][0]/(2*np.pi)**3 - 0.125) < 5e-3
