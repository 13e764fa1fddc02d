"""Multi-GPU parity driver (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_parity.py

Every rank evaluates the RHS of its brick partition on its B200 with the
halo exchange over NCCL and compares it with the NumPy oracle run on the
same partitioning (all ranks in one process, halos delivered between graph
stages).  The partitioned oracle, not the single-partition one, is the
reference: for LDG beta != 0 the reference orients the one-sided fluxes of
inter-partition faces by rank parity (pyfr/solvers/baseadvecdiff/
inters.py:47-58), so partitioning changes the discretisation itself."""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from pyfr_b200 import cases                                  # noqa: E402
from pyfr_b200.backend import B200Backend                    # noqa: E402
from pyfr_b200.comm import NCCLComm                          # noqa: E402
from pyfr_b200.host.system import get_system                 # noqa: E402
from util import oracle_rhs, rel_err                         # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    parts = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    n = tuple(3*p for p in parts)
    ok = True
    comm = None

    for case, kw in [('tgv', dict(order=3, rsolver='hllc')),
                     ('tgv', dict(order=4)), ('tgv', dict(order=2, beta=0.0))]:
        cfg, box = cases.make(case, n, warp=0.1, **kw)
        cfg.set('backend-b200', 'device-id', os.environ.get('LOCAL_RANK', 0))
        be = B200Backend(cfg)
        # One communicator for the whole run (NCCL bring-up takes ~30 s)
        if comm is None:
            comm = NCCLComm(be.rt, rank, world)
        be.comm = comm

        vparts = box.brick_partition(parts)
        mesh = box.local_mesh(vparts, rank)
        sysm = get_system(be, mesh, cfg, 2, comm=comm)

        for _ in range(2):
            sysm.rhs(0.0, 0, 1)
        be.wait()
        out = sysm.ele_scal_upts(1)[0]

        _, ref = oracle_rhs(case, n, warp=0.1, vparts=vparts, nparts=world,
                            **kw)
        _, ext = oracle_rhs(case, n, warp=0.1, vparts=vparts, nparts=world,
                            extended=True, **kw)
        gidx = mesh.eidxs['hex']
        err = rel_err(out, ext[rank])
        floor = rel_err(ref[rank], ext[rank])
        good = err <= max(1e-12, 4*floor)
        ok &= bool(good)
        print(f'[rank {rank}/{world}] {case} {kw}: neles={len(gidx)} '
              f'nbrs={sorted(mesh.con_p)} err={err:.2e} floor={floor:.2e} '
              f'{"PASS" if good else "FAIL"}', flush=True)

    # No ncclCommDestroy: it blocks while captured graphs are alive
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == '__main__':
    main()
