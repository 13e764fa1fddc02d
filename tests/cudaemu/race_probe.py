"""Worker for tests/test_kernel_races.py: runs the fused Navier-Stokes and
Euler RHS through the emulated kernels (built with ThreadSanitizer when
PYFR_B200_EMU_TSAN=1).  ``--drop-barrier`` removes the __syncthreads()
between phases 4 and 5 of gradflux: the negative control."""

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), HERE]

import emu                                                   # noqa: E402

if '--drop-barrier' in sys.argv:
    _orig = emu.translate

    def _bad(src):
        if 'gradflux(' in src:
            i = src.index('// ---- phase 5')
            # (the sum-factorised kernel synchronises through GSYNC())
            bar = 'GSYNC();' if 'GSYNC();' in src[:i] else '__syncthreads();'
            j = src.rindex(bar, 0, i)
            src = src[:j] + src[j + len(bar):]
        return _orig(src)

    emu.translate = _bad

import pyfr_b200.backend as bk                               # noqa: E402
import pyfr_b200.compiler as comp                            # noqa: E402

bk.load_runtime = lambda device=0, dry=False: emu.EmuRuntime()
comp.KernelCompiler.cubin = lambda self, src, name: src.encode()

from pyfr_b200 import cases                                  # noqa: E402
from pyfr_b200.backend import B200Backend                    # noqa: E402
from pyfr_b200.host.system import get_system                 # noqa: E402

runs = [('tgv', (3, 2, 2), dict(order=2, warp=0.1), {}),
        ('tgv', (3, 2, 2), dict(order=2), {}),
        ('tgv', 2, dict(order=4), {}),
        ('vortex', 5, dict(order=3), {}),
        # opt-in variants: 16-byte accesses over column / point pairs
        ('tgv', (3, 2, 2), dict(order=2, warp=0.1),
         {'gradflux-vec2': 'p1,p3,p5', 'conu-pairs': 1,
          'inters-order': 'address'}),
        ('tgv', 2, dict(order=4), {'gradflux-vec2': 'p1,p3,p5'}),
        # the table-driven fused kernel, warp groups, narrow blocks with
        # two CTAs per SM, fp32 (four columns per access)
        ('tgv', (3, 2, 2), dict(order=2, warp=0.1), {'gradflux-tensor': 0}),
        ('tgv', 2, dict(order=4), {'gradflux-groups': 2}),
        ('tgv', (3, 2, 2), dict(order=3), {'n-soa': 4}),
        ('tgv', 2, dict(order=2, precision='single'), {}),
        # persistent loops of several blocks per CTA: the gathered common
        # solution (bulk-copied rows, per-thread copies, indices fetched an
        # iteration ahead, all completing on the block's mbarrier); the
        # per-point form; the opt-in half-block kernel
        ('tgv', (9, 3, 2), dict(order=2, warp=0.1), {'sm-count': 2}),
        ('tgv', (9, 3, 2), dict(order=2), {'sm-count': 2, 'gather-rows': 0}),
        ('tgv', 2, dict(order=4), {'gradflux-split': 1})]
if '--drop-barrier' in sys.argv:
    runs = runs[:1]

for case, n, kw, opts in runs:
    cfg, box = cases.make(case, n, **kw)
    cfg.set('backend-b200', 'graphs', 'false')
    for k, v in opts.items():
        cfg.set('backend-b200', k, v)
    s = get_system(B200Backend(cfg), box.local_mesh(), cfg, 2)
    s.rhs(0.0, 0, 1)
    s.rhs(0.0, 0, 1)

if '--drop-barrier' not in sys.argv:
    # Mixed element types: fused kernels over triangle / prism operators
    for pattern, n, kw in [('quad+tri', (4, 3), dict(order=3)),
                           ('hex+pri', (3, 2, 2), dict(order=2))]:
        cfg, box, _ = cases.mixed_case(pattern, n, **kw)
        cfg.set('backend-b200', 'graphs', 'false')
        s = get_system(B200Backend(cfg), box.local_mesh(), cfg, 2)
        s.rhs(0.0, 0, 1)

    # Two partitions on one device: the element kernel as an interior
    # launch (blocks drawn dynamically from a device counter) and a
    # partition-boundary launch
    from pyfr_b200.comm import LoopbackWorld

    world, systems = LoopbackWorld(2), []
    for r in range(2):
        cfg, box = cases.make('tgv', (64, 2, 2), order=2)
        cfg.set('backend-b200', 'graphs', 'false')
        cfg.set('backend-b200', 'sm-count', 3)
        comm = world.peer(r)
        be = B200Backend(cfg, comm=comm)
        comm.rt = be.rt
        vparts = box.brick_partition((2, 1, 1))
        systems.append(get_system(be, box.local_mesh(vparts, r), cfg, 2,
                                  comm=comm))
    for _ in range(2):
        world.run_lockstep(systems, 0.0, 0, 1)
    assert any(k.info.get('part') == 'interior'
               for w, k in systems[0].rhs_graphs(0, 1)[0].plan
               if w == 'kernel' and getattr(k, 'info', None))

    # Shared-memory tree reduction + atomics (error norm), rkvdh2 stages
    from pyfr_b200.host.integrator import PIController, RK45Stepper

    cfg, box = cases.make('vortex', (4, 4), order=3)
    cfg.set('backend-b200', 'graphs', 'false')
    for k, v in (('dt', 0.05), ('atol', 1e-6), ('rtol', 1e-6)):
        cfg.set('solver-time-integrator', k, v)
    s = get_system(B200Backend(cfg), box.local_mesh(), cfg, 4)
    pi = PIController(RK45Stepper(s, errest=True), cfg,
                      ['rho', 'rhou', 'rhov', 'E'])
    pi.advance_to(0.06)

print('PROBE DONE')
