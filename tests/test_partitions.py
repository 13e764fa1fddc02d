"""The partition-boundary path -- ``pack``, the device-resident exchange
matrices, ``mpiconu`` and ``mpicflux`` -- with several partitions on ONE
device (``pyfr_b200.comm.LoopbackWorld``: one backend per partition, halos
copied device-to-device between the stages of an RHS evaluation), against
the *partitioned* oracle.

Reference semantics: ``pyfr/solvers/base/system.py:185-202`` (exchange
registration), ``pyfr/solvers/baseadvecdiff/inters.py:47-58`` (orientation
of one-sided LDG fluxes on partition faces by rank parity -- a partitioned
run is a different discretisation from the single-partition one unless
``beta = 0``), ``pyfr/backends/cuda/packing.py:34-112``,
``pyfr/backends/base/types.py:250-257``.

Every case runs twice: on the CPU execution model of the generated kernels
(every CPU run) and, marked ``gpu``, on the device through the C ABI with
the RHS graphs captured and replayed as CUDA graphs.
"""

import os
import sys

import numpy as np
import pytest

from pyfr_b200 import cases
from pyfr_b200.comm import LoopbackWorld
from pyfr_b200.host.system import get_system

from util import assert_parity, oracle_rhs, rel_err, rhs_magnitude

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                'cudaemu'))
import emu                                                   # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

SUBSTRATES = ['emulated', pytest.param('device', marks=pytest.mark.gpu)]


def _fixture_parts(name):
    return np.load(os.path.join(GOLDEN, f'{name}.npz'))['vparts']


def _brick(parts):
    return lambda box: box.brick_partition(parts)


# (case, mesh, partitioning, keyword arguments)
#  - bricks in one, two and three directions (1, 3 and 7 neighbours)
#  - irregular partitions made by the reference's BaselinePartitioner
#    (tests/golden/conn_*.npz: uneven neighbour sets, several faces per
#    neighbour pair)
#  - beta in {1/2, 0, -1/2}: which side sends gradients differs
#  - Rusanov and HLLC, Euler and Navier-Stokes, curved and affine elements
CASES = [
    ('tgv', (4, 3, 3), _brick((2, 1, 1)), dict(order=2, warp=0.1)),
    ('tgv', (4, 3, 3), _brick((2, 1, 1)),
     dict(order=3, beta=0.0, rsolver='hllc', warp=0.1)),
    ('tgv', (4, 3, 3), _brick((2, 1, 1)),
     dict(order=2, beta=-0.5, curved=0.5, warp=0.1)),
    ('tgv', (4, 2, 2), _brick((2, 1, 1)), dict(order=4)),
    ('tgv', (4, 4, 2), _brick((2, 2, 1)), dict(order=2, rsolver='hllc')),
    ('tgv', (4, 4, 4), _brick((2, 2, 2)), dict(order=2, warp=0.1)),
    ('tgv', (5, 4, 3), 'conn_hex_periodic_3parts', dict(order=2, warp=0.1)),
    ('tgv', (5, 4, 3), 'conn_hex_periodic_3parts',
     dict(order=3, beta=0.0, rsolver='hllc')),
    ('tgv', (5, 4, 3), 'conn_hex_periodic_3parts',
     dict(order=2, beta=-0.5, warp=0.1)),
    # long enough in x for element blocks without a partition-boundary
    # face: the element kernel runs as an interior launch (first graph,
    # behind the exchange of the traces) and a boundary launch
    ('tgv', (64, 2, 2), _brick((2, 1, 1)), dict(order=2, warp=0.05)),
    ('tgv', (64, 2, 2), _brick((2, 1, 1)), dict(order=2, beta=-0.5)),
    ('vortex', 12, _brick((2, 1)), dict(order=3)),
    ('vortex', 12, _brick((2, 2)), dict(order=2, rsolver='hllc')),
]


def _ids(c):
    case, n, part, kw = c
    pn = part if isinstance(part, str) else 'brick'
    return f'{case}-{n}-{pn}-' + ','.join(f'{k}={v}' for k, v in kw.items())


def _b200_systems(case, n, vparts, nparts, kw, opts, mk=cases.make):
    from pyfr_b200.backend import B200Backend

    world = LoopbackWorld(nparts)
    systems = []
    for r in range(nparts):
        cfg, box = mk(case, n, **kw)[:2]
        for k, v in opts.items():
            cfg.set('backend-b200', k, v)
        comm = world.peer(r)
        be = B200Backend(cfg, comm=comm)
        systems.append(get_system(be, box.local_mesh(vparts, r), cfg, 2,
                                  comm=comm))

    return world, systems


def _kinds(sysm):
    return [getattr(k, 'kind', None) or getattr(getattr(k, 'fn', None),
                                                 'name', None)
            for g in sysm.rhs_graphs(0, 1) for w, k in g.plan
            if w == 'kernel']


@pytest.mark.parametrize('substrate', SUBSTRATES)
@pytest.mark.parametrize('case,n,part,kw', CASES, ids=map(_ids, CASES))
def test_partitioned_rhs_matches_partitioned_oracle(substrate, case, n, part,
                                                    kw, monkeypatch):
    if substrate == 'emulated':
        emu.install(monkeypatch)
        opts = {'graphs': 'false'}
    else:
        import __graft_entry__ as g
        g.build_runtime()
        opts = {'graphs': 'true'}

    _, box = cases.make(case, n, **kw)
    vparts = _fixture_parts(part) if isinstance(part, str) else part(box)
    nparts = int(vparts.max()) + 1

    if n == (64, 2, 2):
        # fewer CTAs than chunks of blocks: the interior launch hands out
        # blocks dynamically (and must re-arm its counter for the replay)
        opts['sm-count'] = 3

    world, systems = _b200_systems(case, n, vparts, nparts, kw, opts)

    # Twice: the second evaluation replays the captured graphs and finds
    # the mailboxes of the first one still in place
    for _ in range(2):
        world.run_lockstep(systems, 0.0, 0, 1)
    for s in systems:
        s.backend.wait()

    osys, ref = oracle_rhs(case, n, vparts=vparts, nparts=nparts, **kw)
    esys, ext = oracle_rhs(case, n, vparts=vparts, nparts=nparts,
                           extended=True, **kw)

    viscous = case == 'tgv'
    for r, s in enumerate(systems):
        out = s.ele_scal_upts(1)[0]
        assert out.shape == ref[r].shape
        assert_parity(out, ref[r], ext[r], 1e-12,
                      mag=rhs_magnitude(esys[r])[0],
                      label=f'{_ids((case, n, part, kw))}[{substrate} '
                            f'rank {r}/{nparts}]')

        kinds = _kinds(s)
        assert 'pack' in kinds and 'mpicflux' in kinds
        assert ('mpiconu' in kinds) == viscous and 'copy' not in kinds

        # the per-neighbour kernels of a graph go out as one launch each
        # (pack: once for the solution, once for the gradients)
        assert kinds.count('mpicflux') <= 2 and kinds.count('pack') <= 2
        assert kinds.count('mpiconu') <= 2

        if n == (64, 2, 2):
            g0, g1, _ = s.rhs_graphs(0, 1)
            parts = [[k.info.get('part') for w, k in g.plan
                      if w == 'kernel' and k.kind == 'gradflux']
                     for g in (g0, g1)]
            assert parts == [['interior'], ['boundary']]
            assert 'intconu' not in kinds
            # ... which follows the exchange of its graph
            assert [w for w, k in g0.plan][-2:] == ['xchg', 'kernel']

    # The partitioned discretisation differs from the single-partition one
    # exactly when the LDG flux is one-sided (a check that the test would
    # notice an exchange that silently did nothing)
    if viscous and nparts == 2:
        _, one = oracle_rhs(case, n, **kw)
        gidx = [box.local_mesh(vparts, r).eidxs['hex'] for r in range(2)]
        d = max(rel_err(systems[r].ele_scal_upts(1)[0], one[0][..., gidx[r]])
                for r in range(2))
        assert (d < 1e-10) == (kw.get('beta', 0.5) == 0.0)


@pytest.mark.parametrize('substrate', SUBSTRATES)
def test_partitioned_walls_irregular(substrate, monkeypatch):
    """Four irregular partitions (reference partitioner) of a box with
    walls in two directions: boundary, interior and inter-partition
    interfaces in every partition; time-dependent boundary data."""
    if substrate == 'emulated':
        emu.install(monkeypatch)
        opts = {'graphs': 'false'}
    else:
        import __graft_entry__ as g
        g.build_runtime()
        opts = {'graphs': 'true'}

    from util import LocalComm, OracleBackend, run_lockstep

    n = (4, 4, 4)
    bcs = {'ylo': 'no-slp-adia-wall', 'yhi': 'char-riem-inv',
           'zlo': 'slp-adia-wall', 'zhi': 'no-slp-isot-wall'}
    kw = dict(order=2, rsolver='hllc')
    vparts = _fixture_parts('conn_hex_walls_4parts')
    mk = lambda case, n, **kw: cases.box_case('navier-stokes', n, bcs, **kw)

    world, systems = _b200_systems(None, n, vparts, 4, kw, opts, mk=mk)
    for _ in range(2):
        world.run_lockstep(systems, 0.25, 0, 1)

    outs = []
    for extended in (False, True):
        lw = LocalComm(0, 4)
        osys = []
        for r in range(4):
            cfg, box, _ = cases.box_case('navier-stokes', n, bcs, **kw)
            cfg.set('backend-oracle', 'extended-mul', extended)
            osys.append(get_system(OracleBackend(cfg),
                                   box.local_mesh(vparts, r), cfg, 2,
                                   comm=lw.peer(r)))
        for s in osys:
            for ks in s._get_kernels(0, 1).values():
                for k in ks:
                    if k.rtnames:
                        k.bind(t=0.25)
        run_lockstep(osys, lw, 0.25, 0, 1)
        outs.append((osys, [s.ele_scal_upts(1)[0] for s in osys]))

    (_, ref), (esys, ext) = outs
    for r, s in enumerate(systems):
        assert_parity(s.ele_scal_upts(1)[0], ref[r], ext[r], 1e-12,
                      mag=rhs_magnitude(esys[r])[0],
                      label=f'walls-4parts[{substrate} rank {r}/4]')
        kinds = _kinds(s)
        assert 'bccflux' in kinds and 'mpicflux' in kinds
