# r01v: fused Euler element kernel (fluxdiv) parity + whole suite regression
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -4
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
from pyfr_b200 import cases
from pyfr_b200.backend import B200Backend
from pyfr_b200.host.system import get_system
for fus in (0, 1):
    cfg, box = cases.make('vortex', 512, order=3)
    cfg.set('backend-b200', 'euler-fusion', fus)
    be = B200Backend(cfg)
    s = get_system(be, box.local_mesh(), cfg, 2)
    for _ in range(5): s.rhs(0.0, 0, 1)
    be.wait(); t = time.perf_counter()
    for _ in range(50): s.rhs(0.0, 0, 1)
    be.wait(); dt = (time.perf_counter() - t)/50
    nd = sum(s.ele_ndofs)
    print(f'euler vortex 512^2 p3 fusion={fus}: {dt*1e3:.3f} ms/RHS, {nd/dt/1e9:.2f} GDoF/s', flush=True)
PY
