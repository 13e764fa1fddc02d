"""Race detection for the generated kernels: the emulated kernels are built
with ThreadSanitizer and the fused Navier-Stokes (general and affine
geometry) and Euler element kernels, the operator and the interface
kernels are run under it.  Every pair of accesses to shared or global
memory by different CUDA threads that is not ordered by a __syncthreads or
an mbarrier shows up as a data race; a deliberately removed barrier must be
reported (negative control), the shipped kernels must be clean."""

import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _libtsan():
    for cc in ('/usr/bin/gcc', 'gcc'):
        try:
            p = subprocess.run([cc, '-print-file-name=libtsan.so'],
                               capture_output=True, text=True).stdout.strip()
        except OSError:
            continue
        if os.path.isabs(p) and os.path.exists(p):
            return p
    return None


def _probe(*args):
    # (BLAS worker threads of NumPy are not instrumented and would show up
    # as false positives: keep NumPy single threaded, and only count
    # reports located in the emulated kernels' own objects)
    env = dict(os.environ, PYFR_B200_EMU_TSAN='1', LD_PRELOAD=_libtsan(),
               TSAN_OPTIONS='halt_on_error=0 exitcode=0',
               OPENBLAS_NUM_THREADS='1', OMP_NUM_THREADS='1')
    res = subprocess.run(
        [sys.executable, os.path.join(HERE, 'cudaemu', 'race_probe.py'),
         *args], capture_output=True, text=True, timeout=1200, env=env
    )
    out = res.stdout + res.stderr
    assert 'PROBE DONE' in out, out[-2000:]
    return sum(1 for l in out.splitlines()
               if l.startswith('SUMMARY: ThreadSanitizer: data race') and
               'pyfr_b200_cudaemu' in l)


@pytest.mark.skipif(_libtsan() is None, reason='libtsan not available')
def test_generated_kernels_are_race_free():
    assert _probe('--drop-barrier') > 0       # the detector sees a real race
    assert _probe() == 0
