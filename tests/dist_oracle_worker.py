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

    flag = torch.tensor([int(ok)])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == '__main__':
    main()
