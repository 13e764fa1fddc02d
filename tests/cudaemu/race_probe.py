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
            j = src.rindex('__syncthreads();', 0, i)
            src = src[:j] + src[j + 16:]
        return _orig(src)

    emu.translate = _bad

import pyfr_b200.backend as bk                               # noqa: E402
import pyfr_b200.compiler as comp                            # noqa: E402

bk.load_runtime = lambda device=0, dry=False: emu.EmuRuntime()
comp.KernelCompiler.cubin = lambda self, src, name: src.encode()

from pyfr_b200 import cases                                  # noqa: E402
from pyfr_b200.backend import B200Backend                    # noqa: E402
from pyfr_b200.host.system import get_system                 # noqa: E402

runs = [('tgv', (3, 2, 2), dict(order=2, warp=0.1)),
        ('tgv', (3, 2, 2), dict(order=2)),
        ('tgv', 2, dict(order=4)),
        ('vortex', 5, dict(order=3))]
if '--drop-barrier' in sys.argv:
    runs = runs[:1]

for case, n, kw in runs:
    cfg, box = cases.make(case, n, **kw)
    cfg.set('backend-b200', 'graphs', 'false')
    s = get_system(B200Backend(cfg), box.local_mesh(), cfg, 2)
    s.rhs(0.0, 0, 1)
    s.rhs(0.0, 0, 1)

print('PROBE DONE')
