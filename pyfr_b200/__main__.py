"""Stand-alone driver: time-step one of the built-in cases on the b200
backend and print the integrate-plugin diagnostics.

    python -m pyfr_b200 tgv --n 32 --order 4 --dt 1e-3 --steps 200 --every 50
    python -m pyfr_b200 vortex --scheme rk45 --atol 1e-6 --rtol 1e-6 \\
        --dt 1e-2 --steps 100 --every 25      # adaptive: PI controller

(Inside a PyFR checkout the backend is used through ``pyfr run -b b200``
instead; this driver exists because PyFR itself is not
installable offline.)  One process per GPU under ``torchrun``-style
launchers (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT).
"""

import argparse
import os
import sys
import time

import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser(prog='python -m pyfr_b200')
    ap.add_argument('case', choices=['tgv', 'vortex'])
    ap.add_argument('--n', type=int, default=16, help='elements per '
                    'direction per rank')
    ap.add_argument('--order', type=int, default=3)
    ap.add_argument('--precision', default='double',
                    choices=['double', 'single'])
    ap.add_argument('--rsolver', default='rusanov',
                    choices=['rusanov', 'hllc'])
    ap.add_argument('--dt', type=float, default=1e-3)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--every', type=int, default=10, help='print the '
                    'integrals every so many steps')
    ap.add_argument('--scheme', default='rk4', choices=['rk4', 'rk45'],
                    help='rk45 runs under the PI step-size controller: '
                    '--dt is the initial step and --steps*--dt the end time')
    ap.add_argument('--cfl', type=float, default=None, help='choose the '
                    'step size from the largest wave speed (CFL '
                    'controller); --steps*--dt is the end time')
    ap.add_argument('--fused-update', action='store_true', help='rk45: '
                    'apply each stage update in the epilogue of the last RHS '
                    'kernel')
    ap.add_argument('--atol', type=float, default=1e-6)
    ap.add_argument('--rtol', type=float, default=1e-6)
    ap.add_argument('--opt', action='append', default=[],
                    help='[backend-b200] option key=value')
    args = ap.parse_args(argv)

    from pyfr_b200 import cases
    from pyfr_b200.backend import B200Backend
    from pyfr_b200.host.integrator import (CFLController, FieldIntegrator,
                                           PIController, RK4Stepper,
                                           RK45Stepper, TGV_EXPRS)
    from pyfr_b200.host.system import get_system

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    nd = 3 if args.case == 'tgv' else 2
    parts = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    parts = parts[:nd]
    if int(np.prod(parts)) != world:
        sys.exit(f'{world} ranks do not tile a {nd}-D box')

    cfg, box = cases.make(args.case, tuple(args.n*p for p in parts),
                          order=args.order, precision=args.precision,
                          rsolver=args.rsolver)
    for kv in args.opt:
        k, v = kv.split('=', 1)
        cfg.set('backend-b200', k, v)

    be = B200Backend(cfg)
    comm = None
    if world > 1:
        from pyfr_b200.comm import NCCLComm
        comm = be.comm = NCCLComm(be.rt, rank, world)

    vparts = box.brick_partition(parts) if world > 1 else None
    adaptive = args.scheme == 'rk45' or args.cfl is not None
    sysm = get_system(be, box.local_mesh(vparts, rank), cfg,
                      4 if args.scheme == 'rk45' else 3, comm=comm,
                      needs_cfl=args.cfl is not None)
    ndof = sum(sysm.ele_ndofs)*world

    exprs = TGV_EXPRS if args.case == 'tgv' else [
        'rho', '0.5*rho*(u*u + v*v)'
    ]
    fi = FieldIntegrator(sysm, cfg, exprs)

    bufs = {}

    def allreduce(vals, op='sum'):
        if world == 1:
            return vals
        arr = np.atleast_1d(np.asarray(vals, dtype=float))
        if len(arr) not in bufs:
            bufs[len(arr)] = be.matrix((1, len(arr)), tags={'noblock'})
            be.commit()
        buf = bufs[len(arr)]
        buf.set(arr[None])
        # (NCCL dtype code of the buffer's precision: 1 = float64)
        comm.allreduce(buf.data, len(arr),
                       1 if be.fpdtype == np.float64 else 0,
                       {'sum': 0, 'max': 2}[op], be.stream)
        be.wait()
        out = buf.get()[0]
        return out if np.ndim(vals) else float(out[0])

    sect = 'solver-time-integrator'
    for k in ('dt', 'atol', 'rtol'):
        cfg.set(sect, k, getattr(args, k))

    if args.cfl is not None:
        cfg.set(sect, 'cfl', args.cfl)
        st = (RK45Stepper(sysm, fused=args.fused_update)
              if args.scheme == 'rk45' else RK4Stepper(sysm))
        ctl = CFLController(st, cfg, allreduce=allreduce)
    elif adaptive:
        convars = ['rho', 'rhou', 'rhov', 'rhow'][:nd + 1] + ['E']
        st = RK45Stepper(sysm, errest=True, fused=args.fused_update)
        ctl = PIController(st, cfg, convars, allreduce=allreduce)
    else:
        st = RK4Stepper(sysm)

    def report():
        vals = allreduce(fi(st.tcurr, st.idxcurr))
        nsteps = ctl.nacptsteps if adaptive else st.nsteps
        if rank == 0:
            print(f'{nsteps:8d} {st.tcurr:12.6f} '
                  + ' '.join(f'{v:.12e}' for v in vals), flush=True)

    if rank == 0:
        print(f'# {args.case}: {ndof} DoF on {world} rank(s); columns: '
              f'step t {" ".join(f"int({e})" for e in exprs)}')
    report()

    be.wait()
    t0 = time.perf_counter()
    for i in range(1, args.steps + 1):
        if not adaptive:
            st.step(args.dt)
        if i % args.every == 0 or i == args.steps:
            if adaptive:
                ctl.advance_to(i*args.dt)
            report()
    be.wait()
    dt = time.perf_counter() - t0

    if rank == 0 and adaptive:
        nrhs = (5 if args.scheme == 'rk45' else 4)*(ctl.nacptsteps
                                                    + ctl.nrjctsteps)
        print(f'# {ctl.nacptsteps} accepted / {ctl.nrjctsteps} rejected '
              f'{args.scheme} '
              f'steps in {dt:.3f} s: {nrhs*ndof/dt/1e9:.3f} GDoF-RHS/s '
              'including the register updates, error norms and diagnostics')
    elif rank == 0:
        print(f'# {args.steps} RK4 steps in {dt:.3f} s: '
              f'{4*args.steps*ndof/dt/1e9:.3f} GDoF-RHS/s including the '
              'register updates and diagnostics')

    if world > 1:
        sys.stdout.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
