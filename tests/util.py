"""Shared helpers for the parity tests."""

import numpy as np

from oracle.npbackend import LocalComm, make_backend
from pyfr_b200 import base, cases
from pyfr_b200.host.system import get_system

OracleBackend = make_backend(base)


def oracle_rhs(case, n, nregs=2, vparts=None, nparts=1, extended=False, **kw):
    """RHS of bank 0 into bank 1 on the NumPy oracle; returns per-rank
    (system, rhs array) lists."""
    world = LocalComm(0, nparts)
    systems = []

    for r in range(nparts):
        cfg, box = cases.make(case, n, **kw)
        cfg.set('backend-oracle', 'extended-mul', extended)
        be = OracleBackend(cfg)
        mesh = box.local_mesh(vparts, r)
        systems.append(get_system(be, mesh, cfg, nregs, comm=world.peer(r)))

    run_lockstep(systems, world, 0.0, 0, 1)
    return systems, [s.ele_scal_upts(1)[0] for s in systems]


def run_lockstep(systems, world, t, uin, fout):
    """Advance all in-process ranks graph by graph, delivering the halo
    messages between stages (what MPI/NCCL do between real ranks)."""
    graphs = [s.rhs_graphs(uin, fout) for s in systems]

    for stage in zip(*graphs):
        for g in stage:
            g.run()
        world.deliver()


def rel_err(a, b):
    return np.abs(a - b).max()/np.abs(b).max()


# Every parity comparison of a test session: (test id, err, floor, ratio).
# tests/conftest.py prints them at the end of the run and writes them to
# gpurun_out/parity_errors.json, so the achieved errors are on record
PARITY_LOG = []

# Largest admissible |out - ext| at a point in units of eps * (magnitude of
# the terms the point's RHS is summed from), see rhs_magnitude()
RUNNING_ERROR_C = 64.0


def _test_id():
    import os
    return os.environ.get('PYTEST_CURRENT_TEST', '?').split(' ')[0]


def rhs_magnitude(sysm, uin=0, fout=1):
    """Per-point magnitude of the terms the RHS in bank ``fout`` of an
    *oracle* system was summed from (a running-error bound in the sense of
    Higham, Accuracy and Stability, section 3.3, carried through the last
    three stages of the path):

        |rcpdjac| ( |M1 - M3 M2| |F|  +  |M3| |n| lambda |u_f| )

    ``F``: the transformed flux the oracle left in ``vect_upts``
    (``pyfr/solvers/baseadvec/elements.py:97-119``); ``lambda |u_f|``: the
    magnitude of the terms of the interface flux -- a Riemann solver forms
    ``(F_L + F_R).n/2 - lambda (u_R - u_L)/2`` with ``lambda = |v| + c``,
    so at low Mach number its *terms* exceed its value by ``c/|v|``
    (``pyfr/solvers/euler/kernels/rsolvers/rusanov.mako``).  A backend that
    evaluates the same sums in another order may differ from the
    extended-precision evaluation by a modest multiple of ``eps`` times
    this quantity at every single point; unlike a tolerance on the field
    maximum that criterion is local: it fails for an error at a point whose
    own terms are small."""
    from pyfr_b200.host.elements import EulerElements, NavierStokesElements
    from pyfr_b200.host.shapes import shape_map

    cfg, mesh = sysm.cfg, sysm.mesh
    cls = {'euler': EulerElements, 'navier-stokes': NavierStokesElements}[
        cfg.get('solver', 'system')]
    gamma = cfg.getfloat('constants', 'gamma')

    mags = []
    for (et, spts), banks in zip(mesh.spts.items(), sysm.ele_banks):
        e = cls(shape_map[et], spts, cfg)
        nd = e.ndims
        bank = banks[fout]

        # Trace of the solution and the magnitude of the Riemann terms
        uf = np.einsum('fu,uvn->fvn', e.basis.opmat('M0'), banks[uin].get())
        rho, E = uf[:, 0], uf[:, -1]
        v2 = sum((uf[:, 1 + d]/rho)**2 for d in range(nd))
        p = (gamma - 1)*(E - 0.5*rho*v2)
        lam = np.sqrt(v2) + np.sqrt(np.abs(gamma*p/rho))
        magn = np.linalg.norm(e._pnorm_fpts, axis=-1)
        mfc = (magn*lam)[:, None, :]*np.abs(uf)

        S = np.zeros(bank.ioshape)
        for A, b, out, alpha, beta in sysm.backend.mul_log:
            if out is not bank:
                continue
            if beta:
                B = mfc
            else:
                B = np.abs(b.get()).reshape(A.shape[1], *bank.ioshape[1:])
            S += abs(alpha)*np.einsum('mk,kvn->mvn', np.abs(A), B)

        mags.append(np.abs(e.rcpdjac_at_np('upts'))[:, None, :]*S)

    return mags


def rhs_magnitude_from_state(cfg, mesh, u, eidx=None):
    """The magnitude field of ``rhs_magnitude`` formed from the *solution*
    alone (no oracle run), for the elements ``eidx`` of a one-element-type
    mesh: the transformed flux is taken as the inviscid one (the viscous
    flux is smaller by the Reynolds number) and the Riemann terms from the
    element's own trace.  What the full-size parity test uses, where no
    extended-precision evaluation is affordable."""
    from pyfr_b200.host.elements import EulerElements, NavierStokesElements
    from pyfr_b200.host.shapes import shape_map

    cls = {'euler': EulerElements, 'navier-stokes': NavierStokesElements}[
        cfg.get('solver', 'system')]
    gamma = cfg.getfloat('constants', 'gamma')
    (et, spts), = mesh.spts.items()
    if eidx is not None:
        spts, u = spts[:, eidx], u[..., eidx]

    e = cls(shape_map[et], spts, cfg)
    nd, nu = e.ndims, e.nupts

    def prim(q):
        rho, E = q[:, 0], q[:, -1]
        vel = [q[:, 1 + d]/rho for d in range(nd)]
        p = (gamma - 1)*(E - 0.5*rho*sum(v*v for v in vel))
        return rho, vel, E, p

    # |S F_inv(u)| at the solution points, term by term
    rho, vel, E, p = prim(u)
    F = np.zeros((nd,) + u.shape)
    for d in range(nd):
        F[d, :, 0] = np.abs(rho*vel[d])
        for i in range(nd):
            F[d, :, 1 + i] = np.abs(rho*vel[d]*vel[i]) + (d == i)*np.abs(p)
        F[d, :, -1] = np.abs((E + p)*vel[d])
    smat = np.abs(e.smat_at_np('upts'))              # (nd, nupts, nd, neles)
    Ft = np.einsum('dpke,kpve->dpve', smat, F)

    mm = lambda A, B: (A @ B.reshape(B.shape[0], -1)).reshape(
        A.shape[0], *B.shape[1:])
    A5 = np.abs(e.basis.opmat('M1 - M3*M2'))
    S = mm(A5, Ft.reshape(nd*nu, *u.shape[1:]))

    uf = mm(e.basis.opmat('M0'), u)
    rho, vel, E, p = prim(uf)
    lam = np.sqrt(sum(v*v for v in vel)) + np.sqrt(np.abs(gamma*p/rho))
    magn = np.linalg.norm(e._pnorm_fpts, axis=-1)
    S += mm(np.abs(e.basis.opmat('M3')), (magn*lam)[:, None, :]*np.abs(uf))

    return np.abs(e.rcpdjac_at_np('upts'))[:, None, :]*S


def geometry_conditioning(mesh, eidx=None):
    """``1 + |x|/h`` per element: how much of an element's metric terms is
    decided by the rounding of its own vertex coordinates.  A linear
    element's Jacobian is a difference of vertex coordinates (``pyfr/
    solvers/baseadvec/kernels/smats.mako``); with coordinates of magnitude
    ``|x|`` and an element of size ``h`` its relative accuracy is ``eps
    |x|/h`` whatever the evaluation order, and every term of the RHS is
    linear in it.  Negligible on the small test meshes (|x|/h <= 2), 32 at
    the corners of the 64^3 box."""
    (et, spts), = mesh.spts.items()
    if eidx is not None:
        spts = spts[:, eidx]
    xmax = np.abs(spts).max(axis=(0, 2))
    h = (spts.max(axis=0) - spts.min(axis=0)).min(axis=-1)
    return 1 + xmax/h


def running_error_ratio(out, ref_ext, mag):
    """max over points of ``|out - ext| / (eps * magnitude)``."""
    eps = np.finfo(np.asarray(out).dtype).eps
    return float(np.max(np.abs(out - ref_ext)/(eps*np.maximum(mag, 1e-300))))


def assert_parity(out, ref64, ref_ext, tol=1e-12, slack=4.0, mag=None,
                  label=None):
    """Per-point RHS parity at the tolerance BASELINE.json states.

    ``ref_ext`` is the oracle with its operator products accumulated in
    extended precision.  At low Mach number the RHS is a small difference
    of large flux terms, so the fp64 oracle itself sits a few 1e-12 (of
    the field maximum) away from ``ref_ext`` purely through summation
    order; a backend is held to ``tol`` or to ``slack`` times that
    intrinsic fp64 noise floor, whichever is larger."""
    floor = rel_err(ref64, ref_ext)
    err = rel_err(out, ref_ext)

    # Point-wise criterion: with the oracle's magnitude field every point
    # must sit within RUNNING_ERROR_C eps of the terms it was summed from
    ratio = rfloor = None
    if mag is not None:
        ratio = running_error_ratio(out, ref_ext, mag)
        rfloor = running_error_ratio(ref64, ref_ext, mag)

    PARITY_LOG.append(dict(test=label or _test_id(), err=float(err),
                           floor=float(floor), ratio=ratio,
                           ratio_oracle=rfloor))

    assert err <= max(tol, slack*floor), (err, floor)
    if ratio is not None:
        assert ratio <= RUNNING_ERROR_C, (ratio, rfloor)
    return err, floor


def conservation_defect(cfg, mesh, rhs):
    """|sum over elements and points of w_p |J| RHS| per variable, relative
    to sum w_p |J| |RHS|.  On a periodic domain the flux-reconstruction RHS
    integrates to zero exactly (equal and opposite common fluxes, exact
    quadrature of the flux divergence), whatever the partitioning, the
    Riemann solver or the LDG parameters -- a size-independent parity
    property that needs no reference evaluation."""
    from pyfr_b200.host.elements import EulerElements, NavierStokesElements
    from pyfr_b200.host.shapes import shape_map

    cls = {'euler': EulerElements, 'navier-stokes': NavierStokesElements}[
        cfg.get('solver', 'system')]
    (et, spts), = mesh.spts.items()
    e = cls(shape_map[et], spts, cfg)

    wj = e.basis.upts_wts[:, None]/e.rcpdjac_at_np('upts')
    tot = np.einsum('pe,pve->v', wj, rhs)
    mag = np.einsum('pe,pve->v', wj, np.abs(rhs))

    return tot, mag
