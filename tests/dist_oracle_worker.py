"""Worker for tests/test_dist_gloo.py (launched with torch.distributed.run,
backend gloo, one process per rank, no GPU).

Each rank builds its partition with the host code, runs the RHS on the
NumPy oracle backend and exchanges the halo messages with real
point-to-point communication between processes (``torch.distributed``
isend/irecv standing in for the NCCL send/recv of the device path).  The
result must be bit-identical to the same partitioning run with all ranks in
one process (``LocalComm``), which is what the GPU multi-rank parity driver
(tests/mgpu_parity.py) uses as its reference."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle.npbackend import _mat3                           # noqa: E402
from pyfr_b200 import cases                                  # noqa: E402
from pyfr_b200.host.system import get_system                 # noqa: E402
from util import OracleBackend, oracle_rhs                   # noqa: E402


class GlooComm:
    def __init__(self):
        self.rank, self.size = dist.get_rank(), dist.get_world_size()
        self.pending = []

    def send_init(self, xm, pid, tag):
        comm = self

        class Send:
            def start(self):
                t = torch.from_numpy(np.ascontiguousarray(_mat3(xm)).copy())
                comm.pending.append((dist.isend(t, pid, tag=tag), None, t))

        return Send()

    def recv_init(self, xm, pid, tag):
        comm = self

        class Recv:
            def start(self):
                t = torch.empty(_mat3(xm).shape, dtype=torch.float64)
                comm.pending.append((dist.irecv(t, pid, tag=tag), xm, t))

        return Recv()

    def deliver(self):
        for req, xm, t in self.pending:
            req.wait()
            if xm is not None:
                _mat3(xm)[:] = t.numpy()

        self.pending.clear()


def adaptive_run(comm):
    """RK45 under the PI controller on a partitioned mesh: the error norm
    and the DoF count are combined across ranks (all_reduce standing in
    for the 1-element ncclAllReduce of the device path), so every rank
    takes the same accept / reject decisions -- and the same ones as the
    unpartitioned run."""
    from pyfr_b200.host.integrator import PIController, RK45Stepper

    rank, world = comm.rank, comm.size
    n, kw, tend = (6, 4), dict(order=3), 0.3
    convars = ['rho', 'rhou', 'rhov', 'E']

    def allreduce(x, op):
        t = torch.tensor([float(x)], dtype=torch.float64)
        dist.all_reduce(t, op={'sum': dist.ReduceOp.SUM,
                               'max': dist.ReduceOp.MAX}[op])
        return type(x)(t.item())

    def run(vparts, comm_, allred):
        cfg, box = cases.make('vortex', n, **kw)
        for k, v in (('dt', 0.08), ('atol', 1e-6), ('rtol', 1e-6)):
            cfg.set('solver-time-integrator', k, v)
        be = OracleBackend(cfg)
        mesh = box.local_mesh(vparts, rank if vparts is not None else 0)
        s = get_system(be, mesh, cfg, 4, comm=comm_)
        if comm_ is not None:
            # the oracle backend leaves message progress to the caller
            be.run_graph = lambda g, wait=False: (g.run(), comm_.deliver())
        pi = PIController(RK45Stepper(s, errest=True), cfg, convars,
                          allreduce=allred)
        pi.advance_to(tend)
        return pi, mesh

    cfg, box = cases.make('vortex', n, **kw)
    vparts = box.brick_partition((world, 1) if world == 2 else (2, 2))
    pp, pmesh = run(vparts, comm, allreduce)
    ps, smesh = run(None, None, None)

    mine = pmesh.eidxs['quad']
    same_hist = [w for _, w, _ in pp.stepinfo] == [w for _, w, _ in ps.stepinfo]
    dts = np.allclose([d for d, *_ in pp.stepinfo],
                      [d for d, *_ in ps.stepinfo], rtol=1e-9, atol=0)
    sol = np.abs(pp.stepper.soln[0] - ps.stepper.soln[0][..., mine]).max()
    ok = (same_hist and dts and sol < 1e-11 and pp.nrjctsteps >= 1
          and pp.gndofs == ps.gndofs and pp.tcurr == ps.tcurr == tend)
    print(f'[rank {rank}] adaptive: decisions={same_hist} dt={dts} '
          f'sol={sol:.1e} ok={ok}', flush=True)
    return ok


def main():
    dist.init_process_group('gloo')
    comm = GlooComm()
    rank, world = comm.rank, comm.size
    parts = {2: (2, 1, 1), 4: (2, 2, 1)}[world]
    ok = True

    for case, n, kw in [('tgv', (4, 2, 2), dict(order=2, warp=0.1)),
                        ('tgv', (4, 4, 2), dict(order=1, beta=0.0,
                                                rsolver='hllc')),
                        ('vortex', (6, 4), dict(order=3))]:
        p = parts[:len(n)] if len(n) == 3 else parts[:2]
        cfg, box = cases.make(case, n, **kw)
        vparts = box.brick_partition(p)

        be = OracleBackend(cfg)
        s = get_system(be, box.local_mesh(vparts, rank), cfg, 2, comm=comm)
        for g in s.rhs_graphs(0, 1):
            g.run()
            comm.deliver()
        out = s.ele_scal_upts(1)[0]

        _, ref = oracle_rhs(case, n, vparts=vparts, nparts=world, **kw)
        same = np.array_equal(out, ref[rank])
        ok &= same
        print(f'[rank {rank}] {case} {kw}: bit-identical={same}', flush=True)

    ok &= adaptive_run(comm)

    flag = torch.tensor([int(ok)])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == '__main__':
    main()
