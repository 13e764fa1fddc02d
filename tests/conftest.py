import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')
    config.addinivalue_line(
        'markers', 'slow: CPU-model cases of opt-in kernel variants that are '
        'not defaults (measured alternatives of round 1); run with '
        'PYFR_B200_SLOW=1')


def pytest_collection_modifyitems(config, items):
    """``gpu`` tests need a device: on a machine without one they are
    skipped (with the runtime's own message) rather than failed, so a plain
    ``pytest tests`` is green wherever it runs.  A library that is missing
    or fails to load is *not* a reason to skip."""
    if not os.environ.get('PYFR_B200_SLOW'):
        skip = pytest.mark.skip(reason='opt-in variant: PYFR_B200_SLOW=1')
        for it in items:
            if it.get_closest_marker('slow'):
                it.add_marker(skip)

    gpu = [it for it in items if it.get_closest_marker('gpu')]
    if not gpu:
        return

    from pyfr_b200.lib import B200NoDevice, load_runtime

    try:
        import __graft_entry__ as g
        g.build_runtime()
        load_runtime(0)
    except B200NoDevice as e:
        skip = pytest.mark.skip(reason=f'no CUDA device: {e}')
        for it in gpu:
            it.add_marker(skip)


@pytest.fixture(scope='session')
def built():
    """Make sure the C-ABI library and the kernel cache exist."""
    import __graft_entry__ as g

    g.build_runtime()
    return g


def pytest_terminal_summary(terminalreporter):
    """Achieved parity errors of the session (tests/util.py: PARITY_LOG):
    a compact table in the pytest output and the full list in
    gpurun_out/parity_errors.json."""
    import json

    try:
        from util import PARITY_LOG, RUNNING_ERROR_C
    except ImportError:
        return
    if not PARITY_LOG:
        return

    out = os.path.join(ROOT, 'gpurun_out')
    try:
        os.makedirs(out, exist_ok=True)
        name = 'parity_errors.json'
        if os.environ.get('PYFR_B200_PARITY_TAG'):
            name = f'parity_errors_{os.environ["PYFR_B200_PARITY_TAG"]}.json'
        with open(os.path.join(out, name), 'w') as f:
            json.dump(PARITY_LOG, f, indent=1)
    except OSError:
        pass

    tr = terminalreporter
    fp = [r for r in PARITY_LOG if r['err'] == r['err']]
    for r in fp:
        if r['floor'] != r['floor']:
            r['floor'] = 0.0
    worst = max(fp, key=lambda r: r['err'])
    rr = [r for r in fp if r.get('ratio') is not None]

    tr.write_sep('=', 'parity (|out - oracle_ext| / max|oracle_ext|)')
    tr.write_line(f'{len(fp)} comparisons; largest err {worst["err"]:.2e} '
                  f'(oracle fp64 floor {worst["floor"]:.2e}) in '
                  f'{worst["test"]}')
    if rr:
        w = max(rr, key=lambda r: r['ratio'])
        own = (f' (oracle fp64 itself {w["ratio_oracle"]:.1f})'
               if w.get('ratio_oracle') is not None else '')
        tr.write_line(f'point-wise running-error ratio (limit '
                      f'{RUNNING_ERROR_C:g} eps): {len(rr)} comparisons, '
                      f'largest {w["ratio"]:.1f}{own} in {w["test"]}')
    for r in sorted(fp, key=lambda r: -r['err'])[:6]:
        tr.write_line(f'  err {r["err"]:.2e} floor {r["floor"]:.2e} '
                      + (f'ratio {r["ratio"]:.1f} ' if r.get('ratio')
                         is not None else '') + r['test'].split('::')[-1])
