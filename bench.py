#!/usr/bin/env python
"""RHS throughput benchmark (BASELINE.json metric: GDoF-RHS/s, TGV hex p=4
fp64).

One *step* is one right-hand-side evaluation ``system.rhs(t, 0, 1)`` over
the whole mesh (the metric the reference defines as ``rhs-gdof/s =
gndofs*nrhsevals/wtime``, pyfr/integrators/base.py:348-349).  At N GPUs the
mesh is N bricks of ``--n``^3 hexes (weak scaling), one rank per GPU, halo
exchange over NCCL.

Prints ONE JSON line (rank 0).  ``value`` is timed with CUDA events on the
compute stream with all data resident in HBM; ``e2e`` repeats the
measurement with the solution uploaded from / the RHS downloaded to pinned
host memory inside the timed region; ``roofline`` describes the dominant
kernel (per-launch CUDA-event times against its algorithmic bytes);
``cpu_baseline`` is the C/OpenMP port of the reference's CPU design
(oracle/crhs) timed on the host cores on a bounded sample.  ``--impl reference`` times only that CPU port.
"""

import argparse
import ctypes as ct
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np                                       # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    # (--mesh-n: the spelling to use under torchrun, whose own parser
    # claims --n as an abbreviation of --nnodes)
    ap.add_argument('--n', '--mesh-n', type=int, default=64,
                    help='hexes per direction: '
                    'per GPU (weak scaling) or of the whole mesh (strong)')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: N bricks of n^3 (the default, what the '
                    'driver times); strong: one n^3 mesh split over N ranks '
                    '(BASELINE configs[2]: --scaling strong --n 128)')
    ap.add_argument('--partition', default='brick',
                    choices=['brick', 'reference'],
                    help='brick: equal bricks; reference: the partition the '
                    'reference partitioner (pyfr/partitioners/baseline.py) '
                    'made for this mesh, from tests/golden/parts_hex<n>.npz')
    ap.add_argument('--no-parity', action='store_true', help='skip the '
                    'small partitioned-oracle check that precedes the timing')
    ap.add_argument('--timestep', action='store_true', help='also time whole '
                    'time steps (RK4 over axnpby; RK45 with separate and '
                    'with fused stage updates) and report them as '
                    '"time_step"')
    ap.add_argument('--case', default='tgv',
                    choices=['tgv', 'hex+pri', 'hex+pri+pyr+tet'],
                    help='tgv: the headline workload; the others time '
                    'BASELINE configs[3]-style mixed meshes of --n^3 cells per '
                    'GPU (device-resident figure only)')
    ap.add_argument('--order', type=int, default=4)
    ap.add_argument('--precision', default='double')
    ap.add_argument('--rsolver', default='rusanov')
    ap.add_argument('--cpu-n', type=int, default=24, help='mesh size of the '
                    'bounded CPU-baseline sample')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-graphs', action='store_true')
    ap.add_argument('--no-clocks', action='store_true', help='skip the '
                    'nvidia-smi sampling load (for runs under ncu)')
    ap.add_argument('--opt', action='append', default=[],
                    help='backend option key=value ([backend-b200])')
    ap.add_argument('--kernel-times', default=None, help='write the '
                    'per-kernel event timings to this JSON file')
    return ap.parse_args()


# -- helpers ------------------------------------------------------------------
def bricks(nparts):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[nparts]


class ClockSampler:
    """nvidia-smi clock / throttle sampling while the GPU is under the
    benchmark's load (the profiling recipe's clocks line).  Rows are
    time-stamped on arrival; ``stop`` summarises those that fall inside
    the load window opened by ``mark``."""

    Q = ('index,clocks.sm,clocks.max.sm,power.draw,'
         'clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu, period_ms=50):
        self.gpu, self.rows, self.proc = gpu, [], None
        self.period_ms, self.t0 = period_ms, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}',
                 '--format=csv,noheader,nounits', '-lms', str(self.period_ms),
                 '-i', str(self.gpu)], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True
            )
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(),
                              [c.strip() for c in line.split(',')]))

    def mark(self):
        self.t0 = time.time()

    def nsamples(self):
        return sum(1 for t, r in self.rows if self.t0 and t >= self.t0)

    def stop(self):
        if self.proc is None:
            return None

        t1 = time.time()
        time.sleep(0.1)
        self.proc.terminate()
        self.t.join(timeout=2)

        sm, smax, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for t, r in self.rows:
            if self.t0 is None or not (self.t0 <= t <= t1 + 0.05):
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)

        if not sm:
            return None

        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(smax),
                'reasons': sorted(reasons), 'samples': len(sm),
                'power_w_max': max(pw)}


def cpu_baseline(args, min_seconds=10.0, steps=None, warmup=1, n=None,
                 budget_s=None):
    """The C/OpenMP restatement of the reference's CPU design
    (oracle/crhs, driven through the oracle backend API) on all host
    cores: the same TGV case on an ``n``^3 mesh (default: the bounded
    sample ``--cpu-n``).  Rebuilt on the machine it runs on
    (-march=native); threads pinned (OMP_PROC_BIND=close) and the
    *median* step time reported, because the mean of a few steps on a
    shared host moved by 1.6x between runs."""
    # (libgomp reads these when it is loaded, i.e. with the library below)
    os.environ.setdefault('OMP_PROC_BIND', 'close')
    os.environ.setdefault('OMP_PLACES', 'cores')

    from oracle import cbackend
    from pyfr_b200 import base, cases
    from pyfr_b200.host.system import get_system

    subprocess.run(['make', '-s', '-B', '-C', os.path.join(ROOT, 'oracle')],
                   check=True, capture_output=True)

    n = n or args.cpu_n
    cfg, box = cases.make('tgv', n, order=args.order, rsolver=args.rsolver)
    be = cbackend.make_cbackend(base, fast=True,
                                nthreads=os.cpu_count())(cfg)
    sysm = get_system(be, box.local_mesh(), cfg, 2)
    ndof = sum(sysm.ele_ndofs)

    t0 = time.perf_counter()
    for _ in range(max(warmup, 1)):
        sysm.rhs(0.0, 0, 1)

    # (exactly ``steps`` timed evaluations unless that would exceed the
    # time budget, judged by the warm-up)
    if steps and budget_s:
        est = (time.perf_counter() - t0)/max(warmup, 1)
        steps = max(1, min(steps, int(budget_s/max(est, 1e-9))))

    times, t0 = [], time.perf_counter()
    while (len(times) < steps if steps else
           (len(times) < 3 or time.perf_counter() - t0 < min_seconds)):
        t1 = time.perf_counter()
        sysm.rhs(0.0, 0, 1)
        times.append(time.perf_counter() - t1)
    dt = statistics.median(times)

    return {
        'value': ndof/dt/1e9, 'unit': 'GDoF/s', 'cores': be.nthreads,
        'kind': 'port', 'mesh': f'{n}^3', 'nsteps': len(times),
        'step_s': {'median': dt, 'min': min(times), 'max': max(times)},
        'sample': f'median of {len(times)} RHS evaluations of TGV NS hex '
                  f'p={args.order} fp64 on {n}^3 elements ({ndof} DoF); '
                  'C11/OpenMP restatement of the reference OpenMP backend '
                  'design (blocked AoSoA, block-group fusion with '
                  'thread-local scratch, CSR operator kernels instead of '
                  'libxsmm; gcc -O3 -march=native -ffast-math), all host '
                  'threads of ONE process, pinned; the reference backend '
                  'itself cannot run offline'
    }, dt


def reference_arm(args):
    """``--impl reference``: the CPU implementation of the path alone, on
    the host cores, same metric/config keys; each step is one RHS of the
    bounded sample."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return

    # The mesh of the b200 arm itself when one evaluation stays within a
    # second or so of CPU time (64^3: ~0.5 s), else the bounded sample;
    # at most ~60 s of timed work whatever --steps says
    n = args.n if args.n <= 64 else args.cpu_n
    info, dt = cpu_baseline(args, steps=args.steps,
                            warmup=max(args.warmup, 1), n=n, budget_s=90.0)
    steps = info.pop('nsteps')

    cfgd = workload_config(args, n=n, reference=True)
    v = info['value']
    line = {
        'impl': 'reference', 'metric': 'GDoF-RHS/s', 'value': v,
        'unit': 'GDoF/s', 'n_gpus': args.gpus, 'steps': steps,
        'warmup': args.warmup, 'ms_per_step': dt*1e3,
        'higher_is_better': True, 'scaling': args.scaling,
        'vs_baseline': None,
        'dtype': 'f64' if args.precision == 'double' else 'f32',
        'data': 'synthetic',
        'config': cfgd, 'cpu_baseline': info,
        'e2e': {'value': v, 'unit': 'GDoF/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0}
    }
    print(json.dumps(line))


def workload_config(args, n=None, reference=False):
    n = n or args.n
    if reference:
        # One host process whatever --gpus says: the CPU port has no
        # multi-process driver, so at N > 1 this is a single-host figure
        per = ' per GPU' if n == args.n and args.scaling == 'weak' else ''
        return {
            'workload': f'TGV compressible Navier-Stokes, {n}^3 periodic '
                        f'hexes{per}, p={args.order}, fp64, {args.rsolver}, '
                        'LDG beta=0.5 tau=0.1, one RHS evaluation per step',
            'mesh': f'{n}^3 on one host process (the b200 arm: '
                    f'{args.n}^3 per GPU)' if args.scaling == 'weak' else
                    f'{n}^3 on one host process',
            'l2': 'n/a (CPU)',
            'parallelism': 'OpenMP threads of one process; not partitioned'
                           + (f' (single-host figure beside {args.gpus} '
                              'GPU ranks)' if args.gpus > 1 else '')
        }

    if args.case != 'tgv':
        return {
            'workload': f'compressible Navier-Stokes, {args.n}^3 periodic '
                        f'cells of mixed type ({args.case}), p={args.order}, '
                        f'{"fp64" if args.precision == "double" else "fp32"}'
                        f', {args.rsolver}, one RHS evaluation per step',
            'mesh': f'{args.n}^3 cells per GPU, brick partition',
            'l2': 'inputs larger than L2 for n >= 32',
            'parallelism': f'domain decomposition, {args.gpus} rank(s), NCCL '
                           'send/recv halo exchange'
        }

    strong = args.scaling == 'strong'
    part = ('brick partition' if args.partition == 'brick' else
            "partitioned by the reference's BaselinePartitioner "
            '(tests/golden/parts_hex*.npz)')
    nloc = args.n**3//(args.gpus if strong else 1)
    return {
        'workload': f'TGV compressible Navier-Stokes, {args.n}^3 periodic '
                    f'hexes {"in total" if strong else "per GPU"}, '
                    f'p={args.order}, '
                    f'{"fp64" if args.precision == "double" else "fp32"}, '
                    f'{args.rsolver}, LDG beta=0.5 tau=0.1, one RHS '
                    'evaluation per step',
        'mesh': (f'{args.n}^3 over {args.gpus} GPU(s), {part}' if strong
                 else f'{args.n}^3 per GPU, {part}'),
        'l2': 'inputs larger than L2 (solution bank alone is '
              f'{nloc*(args.order + 1)**3*5*8/1e6:.0f} MB per GPU)',
        'parallelism': f'domain decomposition, {args.gpus} rank(s), NCCL '
                       'send/recv halo exchange'
    }


def reference_partition(nglob, nparts):
    """Element -> rank map made by the reference's own partitioner for this
    box (``pyfr/partitioners/baseline.py``, run offline by
    ``tests/golden/make_golden.py --partitions``; the dual graph of a
    periodic box only depends on its size)."""
    if len(set(nglob)) != 1:
        raise SystemExit('--partition reference needs a cubic mesh')

    path = os.path.join(ROOT, 'tests', 'golden', f'parts_hex{nglob[0]}.npz')
    try:
        with np.load(path) as f:
            return f[f'vparts{nparts}'].astype(np.int32)
    except (OSError, KeyError):
        raise SystemExit(f'No reference partition of a {nglob[0]}^3 box into '
                         f'{nparts} parts ({path}); use --partition brick')


def parity_check(args, be, comm, rank, world):
    """The RHS of a small partitioned TGV mesh (3^3 hexes per rank, same
    order, Riemann solver and exchange path as the timed run) on this
    rank's device against the *partitioned* oracle, before timing.  The
    oracle is used as the checker only.  Returns ``{err, floor, nranks}``
    (maximum over ranks, relative to the field maximum; ``floor`` is the
    oracle's own fp64 distance from its extended-precision evaluation)."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from pyfr_b200 import cases
    from pyfr_b200.host.system import get_system
    from util import oracle_rhs, rel_err

    parts = bricks(world)
    n = tuple(3*p for p in parts)
    kw = dict(order=args.order, rsolver=args.rsolver, warp=0.1)

    cfg, box = cases.make('tgv', n, precision=args.precision, **kw)
    vparts = box.brick_partition(parts) if world > 1 else None
    sysm = get_system(be, box.local_mesh(vparts, rank), cfg, 2, comm=comm)
    for _ in range(2):
        sysm.rhs(0.0, 0, 1)
    be.wait()
    out = sysm.ele_scal_upts(1)[0].astype(float)

    # (floor: the oracle in the working precision against the fp64 oracle
    # with extended-precision operator products)
    _, ref = oracle_rhs('tgv', n, vparts=vparts, nparts=world,
                        precision=args.precision, **kw)
    _, ext = oracle_rhs('tgv', n, vparts=vparts, nparts=world, extended=True,
                        **kw)
    err = rel_err(out, ext[rank])
    floor = rel_err(ref[rank].astype(float), ext[rank])

    if world > 1:
        red = be.matrix((1, 4), tags={'noblock'})
        code = 1 if be.fpdtype == np.float64 else 0
        red.set(np.array([[err, floor, 0.0, 0.0]]))
        comm.allreduce(red.data, 4, code, 2, be.stream)
        be.wait()
        err, floor = (float(x) for x in red.get()[0, :2])

    del sysm
    return {'err': float(err), 'floor': float(floor), 'nranks': world,
            'mesh': 'x'.join(map(str, n)) + f' hexes, p={args.order}, '
            f'{"brick partitions, NCCL halo exchange" if world > 1 else "one partition"}'}


# -- main benchmark -----------------------------------------------------------
def main():
    args = parse()

    if args.impl == 'reference':
        reference_arm(args)
        return

    from pyfr_b200 import cases
    from pyfr_b200.backend import B200Backend
    from pyfr_b200.comm import NCCLComm
    from pyfr_b200.host.system import get_system

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    lrank = int(os.environ.get('LOCAL_RANK', 0))

    if world != args.gpus:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}; launch '
                         'with torch.distributed.run for N > 1')

    # Mesh: `world` bricks of n^3 hexes (weak) or one n^3 mesh (strong)
    parts = bricks(world)
    strong = args.scaling == 'strong'
    nglob = ((args.n,)*3 if strong else tuple(args.n*p for p in parts))
    if strong and any(args.n % p for p in parts):
        raise SystemExit('--scaling strong: n must divide into the bricks')

    if args.case == 'tgv':
        cfg, box = cases.make('tgv', nglob, order=args.order,
                              precision=args.precision, rsolver=args.rsolver)
    else:
        args.no_e2e = args.no_cpu = True
        cfg, box, _ = cases.mixed_case(args.case, nglob, order=args.order,
                                       precision=args.precision,
                                       rsolver=args.rsolver)
    cfg.set('backend-b200', 'device-id', lrank)

    # 64-bit view indices where a rank's largest buffer -- the gradients at
    # the flux points, an allocation of its own -- outgrows what the normal
    # memory model may allocate at once (4*2^31 bytes, as in the reference:
    # [backend] memory-model = large): 128^3, p = 4 on up to four ranks.
    # (64^3 per GPU, the headline configuration, stays at 32-bit indices.)
    nele_loc = int(np.prod(nglob))//world
    n1 = args.order + 1
    isz0 = 8 if args.precision == 'double' else 4
    if args.case == 'tgv' and 3*6*n1**2*5*isz0*nele_loc*1.02 >= 4*2**31:
        cfg.set('backend', 'memory-model', 'large')
    if args.no_graphs:
        cfg.set('backend-b200', 'graphs', 'false')
    for kv in args.opt:
        k, v = kv.split('=', 1)
        cfg.set('backend-b200', k, v)

    be = B200Backend(cfg)
    rt = be.rt

    comm = None
    if world > 1:
        comm = be.comm = NCCLComm(rt, rank, world)
    else:
        comm = type('Serial', (), {'rank': 0, 'size': 1})()

    vparts = None
    if world > 1 and args.partition == 'brick':
        vparts = box.brick_partition(parts)
    elif world > 1:
        vparts = reference_partition(nglob, world)

    # Small partitioned run against the oracle before anything is timed
    parity = None
    if not args.no_parity and args.case == 'tgv':
        parity = parity_check(args, be, comm, rank, world)

    t0 = time.time()
    nregs = 2 if args.no_e2e else 4
    if args.timestep:
        nregs = max(nregs, 3)
    sysm = get_system(be, box.local_mesh(vparts, rank), cfg, nregs, comm=comm)
    setup_s = time.time() - t0
    ndof_local = sum(sysm.ele_ndofs)
    ndof = ndof_local*world
    isz = 8 if args.precision == 'double' else 4

    # Scalar used for barriers / max-over-ranks
    red = be.matrix((1, 4), tags={'noblock'})

    def barrier():
        if world > 1:
            comm.allreduce(red.data, 1, 1 if isz == 8 else 0, 2, be.stream)
        rt.stream_sync(be.stream)

    def max_over_ranks(x):
        if world == 1:
            return x
        red.set(np.full((1, 4), x))
        comm.allreduce(red.data, 1, 1 if isz == 8 else 0, 2, be.stream)
        rt.stream_sync(be.stream)
        return float(red.get()[0, 0])

    ev0, ev1 = rt.new_ptr(rt.event_create), rt.new_ptr(rt.event_create)

    def timed(fn, steps):
        barrier()
        rt.device_sync()
        rt.event_record(ev0, be.stream)
        for _ in range(steps):
            fn()
        rt.event_record(ev1, be.stream)
        rt.event_sync(ev1)
        rt.device_sync()
        ms = rt.elapsed_ms(ev0, ev1)
        barrier()
        return max_over_ranks(ms)

    step = lambda: sysm.rhs(0.0, 0, 1)

    # ---- device-resident throughput --------------------------------------
    sampler = ClockSampler(lrank)
    if rank == 0 and not args.no_clocks:
        sampler.start()

    for _ in range(max(args.warmup, 3)):
        step()
    rt.device_sync()

    sampler.mark()
    l0 = be.nlaunches
    ms = timed(step, args.steps)
    graphs = sysm.rhs_graphs(0, 1)
    nkern = sum(1 for g in graphs for w, k in g.plan if w == 'kernel')
    launches = (be.nlaunches - l0) if not be.use_graphs else nkern*args.steps

    # nvidia-smi needs a second or so of load to return a handful of
    # samples: keep the identical load running (untimed; the same count on
    # every rank, derived from the max-over-ranks time) until it has them
    extra = (0 if args.no_clocks else
             int(np.ceil(max(0.0, 1500.0 - ms)/(ms/args.steps))))
    for _ in range(extra):
        step()
    rt.device_sync()

    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks['load'] = (f'{args.steps} timed + {extra} further identical '
                          'untimed steps')
    ms_per_step = ms/args.steps
    value = ndof/(ms_per_step*1e-3)/1e9

    # ---- per-kernel event timing (dominant kernel + roofline) -------------
    kt = {}
    kernels = [(gi, i, k) for gi, g in enumerate(graphs)
               for i, (w, k) in enumerate(g.plan) if w == 'kernel']
    evs = [(rt.new_ptr(rt.event_create), rt.new_ptr(rt.event_create))
           for _ in kernels]
    nrep = min(args.steps, 10)
    acc = [0.0]*len(kernels)

    if True:
        for _ in range(nrep):
            for (a, b), (gi, i, k) in zip(evs, kernels):
                rt.event_record(a, be.stream)
                k.run(be.stream)
                rt.event_record(b, be.stream)
            rt.device_sync()
            for j, (a, b) in enumerate(evs):
                acc[j] += rt.elapsed_ms(a, b)

        for j, (gi, i, k) in enumerate(kernels):
            name = getattr(getattr(k, 'fn', None), 'name',
                           type(k).__name__)
            meta = (k.misc[0] if getattr(k, 'misc', None) and
                    isinstance(k.misc[0], dict) else {})
            label = f'g{gi}.{i}:{getattr(k, "kind", None) or name}' + (
                f'[{meta["M"]}x{meta["K"]}]' if 'M' in meta else '')
            kt[label] = {'ms': acc[j]/nrep,
                         'bytes': getattr(k, 'traffic', 0)}

    roof = None
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak_gbs = peaks.get('hbm_gbs', 6650.0)
    peak_src = 'measured' if 'hbm_gbs' in peaks else 'fallback'

    # DRAM traffic per launch from the committed ncu --set full capture of
    # this same workload (profiles/ncu_traffic.json), when there is one
    # (a *static* figure: ncu cannot run inside a timed benchmark; valid
    # for the per-GPU workload it was captured on, whatever --gpus is)
    traffic, traffic_src = {}, None
    nloc = args.n if args.scaling == 'weak' else None
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            tj = json.load(f)
        if args.case == 'tgv' and \
           tj['workload'] == (f'tgv n={nloc} order={args.order} '
                              f'{args.precision} {args.rsolver}'):
            traffic = {k: v['dram_bytes'] for k, v in tj['kernels'].items()}
            traffic_src = ('static: profiles/ncu_traffic.json, dram bytes '
                           'per launch from the ncu --set full capture '
                           f'{tj.get("source")} of this per-GPU workload')
    except (OSError, KeyError, ValueError):
        pass

    if kt:
        dom = max(kt, key=lambda n: kt[n]['ms'])
        d = kt[dom]
        ach = d['bytes']/(d['ms']*1e-3)/1e9
        ksum = sum(v['ms'] for v in kt.values())
        roof = {
            'bound': 'hbm', 'kernel': dom, 'achieved': ach, 'peak': peak_gbs,
            'unit': 'GB/s', 'frac': ach/peak_gbs,
            'traffic': traffic.get(dom),
            'traffic_source': traffic_src if dom in traffic else None,
            'algorithmic_bytes': d['bytes'],
            'peak_source': peak_src, 'kernel_ms': d['ms'],
            'kernel_share_of_step': d['ms']/ksum,
            'sum_kernel_ms': ksum
        }

    # Whole-RHS figure against the 3-pass algorithmic model
    nu, nf = (args.order + 1)**3, 6*(args.order + 1)**2
    balg = (5*nu*5 + 5*nf*5 + 3*3*nf*5)*isz/(nu*5)
    rhs_model = {
        'bytes_per_dof_3pass': balg,
        # (per GPU: the whole-job value over the ranks)
        'achieved_gbs_3pass': value*balg/world,
        'frac_of_hbm_3pass': value*balg/world/peak_gbs,
        'peak_source': peak_src
    } if args.case == 'tgv' else None

    # ---- end to end: host buffers in, host buffers out ---------------------
    # Every step uploads its solution from pinned host memory, evaluates
    # the RHS and downloads the result to pinned host memory.  Steps are
    # software-pipelined over two bank pairs (0->1, 2->3) and three streams
    # (upload / compute / download): step i+1's upload and step i-1's
    # download overlap step i's RHS, each ordered by events, so the
    # steady-state step time is the slowest of the three legs (PCIe).
    e2e, e2e_error = None, None

    def measure_e2e():
        banks = sysm.ele_banks[0]
        nb = banks[0].nbytes
        npair = 2 if len(banks) >= 4 else 1
        hin = [rt.new_ptr(rt.malloc_host, nb) for _ in range(npair)]
        hout = [rt.new_ptr(rt.malloc_host, nb) for _ in range(npair)]
        for h in hin:
            rt.memcpy(h, banks[0].data, nb)

        s_in, s_out = (rt.new_ptr(rt.stream_create) for _ in range(2))
        evn = lambda: [rt.new_ptr(rt.event_create) for _ in range(npair)]
        ev_up, ev_rhs, ev_down = evn(), evn(), evn()
        for p in range(npair):
            rt.event_record(ev_rhs[p], be.stream)
            rt.event_record(ev_down[p], be.stream)
        ctr = [0]

        def step_e2e():
            p = ctr[0] % npair
            ctr[0] += 1
            uin, fout = 2*p, 2*p + 1

            # upload once the RHS that last read this input bank is done
            rt.stream_wait_event(s_in, ev_rhs[p])
            banks[uin].upload_packed(hin[p], s_in)
            rt.event_record(ev_up[p], s_in)

            # RHS after its upload and after the previous download of fout
            rt.stream_wait_event(be.stream, ev_up[p])
            rt.stream_wait_event(be.stream, ev_down[p])
            sysm.rhs(0.0, uin, fout)
            rt.event_record(ev_rhs[p], be.stream)

            rt.stream_wait_event(s_out, ev_rhs[p])
            banks[fout].download_packed(hout[p], s_out)
            rt.event_record(ev_down[p], s_out)

        def run_e2e(n):
            # nothing may start before the start event ...
            rt.event_record(ev_up[0], be.stream)
            rt.stream_wait_event(s_in, ev_up[0])
            for _ in range(n):
                step_e2e()
            # ... and the stop event follows the last downloads
            for p in range(npair):
                rt.stream_wait_event(be.stream, ev_down[p])

        run_e2e(4)
        rt.device_sync()

        esteps = max(4, min(args.steps, 10))
        ems = timed(lambda: run_e2e(esteps), 1)/esteps
        res = {'value': ndof/(ems*1e-3)/1e9, 'unit': 'GDoF/s',
               'h2d_bytes_per_step': nb*world, 'd2h_bytes_per_step': nb*world,
               'ms_per_step': ems,
               'api': 'Matrix.upload_packed -> system.rhs -> '
                      'Matrix.download_packed, pinned host memory; steps '
                      f'pipelined over {npair} bank pair(s) on separate '
                      'upload/compute/download streams'}
        for h in hin + hout:
            rt.free_host(h)
        return res

    if not args.no_e2e:
        # On one rank a failure (e.g. pinned host memory refused) is
        # reported in the line instead of losing the whole run; with
        # several ranks it is fatal, because ranks must not diverge
        try:
            e2e = measure_e2e()
        except Exception as exc:                        # pragma: no cover
            if world > 1:
                raise
            e2e_error = f'{type(exc).__name__}: {exc}'

    # ---- whole time steps (the callers of the RHS) ----------
    tstep = None
    if args.timestep:
        from pyfr_b200.host.integrator import RK4Stepper, RK45Stepper

        tstep, dt = {}, 1e-5
        for name, mk, nst in [
                ('rk4', lambda: RK4Stepper(sysm), 4),
                ('rk45', lambda: RK45Stepper(sysm), 5),
                ('rk45_fused_update', lambda: RK45Stepper(sysm, fused=True),
                 5)]:
            st = mk()
            st.advance(3, dt)
            nsteps = max(3, args.steps//4)
            tms = timed(lambda: st.advance(1, dt), nsteps)/nsteps
            tstep[name] = {
                'ms_per_step': tms, 'rhs_per_step': nst,
                'gdof_rhs_per_s': ndof*nst/(tms*1e-3)/1e9,
                'overhead_vs_rhs_only': tms/(nst*ms_per_step) - 1
            }

            if name == 'rk45_fused_update':
                # per-kernel event times of one fused stage (graphs of a
                # key that carries a post-RHS kernel)
                key = next(k for k in sysm._graphs if len(k) == 3)
                ks = [k for g in sysm._graphs[key] for w, k in g.plan
                      if w == 'kernel']
                ea, eb = (rt.new_ptr(rt.event_create) for _ in range(2))
                kms = {}
                for j, k in enumerate(ks):
                    rt.event_record(ea, be.stream)
                    for _ in range(3):
                        k.run(be.stream)
                    rt.event_record(eb, be.stream)
                    rt.device_sync()
                    kms[f'{j}:{getattr(k, "kind", None) or k.fn.name}'] = \
                        rt.elapsed_ms(ea, eb)/3
                tstep[name]['stage_kernels_ms'] = kms

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, _ = cpu_baseline(args)
        cpu.pop('nsteps', None)

    if rank == 0:
        line = {
            'metric': 'GDoF-RHS/s', 'value': value, 'unit': 'GDoF/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': args.scaling, 'vs_baseline': None,
            'dtype': 'f64' if isz == 8 else 'f32', 'data': 'synthetic',
            'config': workload_config(args), 'parity': parity,
            'gpu_launches': launches,
            'launches_per_step': nkern, 'cuda_graphs': be.use_graphs,
            'index_bits': 64 if be.ixdtype == np.int64 else 32,
            'dof': ndof, 'setup_s': setup_s, 'clocks': clocks,
            'roofline': roof, 'rhs_model': rhs_model, 'time_step': tstep,
            'e2e': e2e,
            **({'e2e_error': e2e_error} if e2e_error else {}),
            'cpu_baseline': cpu,
            'compiler': be.compiler.stats
        }
        print(json.dumps(line))

        if args.kernel_times:
            with open(args.kernel_times, 'w') as f:
                json.dump({'ms_per_step': ms_per_step, 'kernels': kt}, f,
                          indent=1)

    if world > 1:
        # Leave the communicator to process exit (ncclCommDestroy blocks
        # while the captured RHS graphs are alive)
        barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
