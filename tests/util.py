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


def assert_parity(out, ref64, ref_ext, tol=1e-12, slack=4.0):
    """Per-point RHS parity at the tolerance BASELINE.json states.

    ``ref_ext`` is the oracle with its operator products accumulated in
    extended precision.  At low Mach number the RHS is a small difference
    of large flux terms, so the fp64 oracle itself sits a few 1e-12 (of
    the field maximum) away from ``ref_ext`` purely through summation
    order; a backend is held to ``tol`` or to ``slack`` times that
    intrinsic fp64 noise floor, whichever is larger."""
    floor = rel_err(ref64, ref_ext)
    err = rel_err(out, ref_ext)

    assert err <= max(tol, slack*floor), (err, floor)
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
