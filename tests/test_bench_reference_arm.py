"""bench.py --impl reference: the CPU arm the driver runs beside the GPU
arm (same metric / config line, host cores only, rank 0 alone under
torchrun)."""

import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = {'impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup',
        'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype',
        'data', 'config', 'cpu_baseline', 'e2e'}


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _check(stdout, ngpus):
    lines = [l for l in stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, stdout[-1500:]
    line = json.loads(lines[0])

    assert KEYS <= set(line)
    assert line['impl'] == 'reference' and line['n_gpus'] == ngpus
    assert line['metric'] == 'GDoF-RHS/s' and line['unit'] == 'GDoF/s'
    assert line['higher_is_better'] is True and line['vs_baseline'] is None
    assert line['value'] > 0 and line['steps'] == 2 and line['warmup'] == 1
    assert 'workload' in line['config']

    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['sample']
    assert cb['value'] == line['value']
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'],
                           'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_single(built):
    res = subprocess.run(
        [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl',
         'reference', '--steps', '2', '--warmup', '1', '--n', '4'],
        capture_output=True, text=True, timeout=600, cwd=ROOT
    )
    assert res.returncode == 0, res.stderr[-2000:]
    _check(res.stdout, 1)


def test_reference_arm_under_torchrun(built):
    """Launched like the scaling runs: rank 0 alone works and prints, the
    other rank exits 0 without output."""
    res = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
         '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
         '--master-port', str(_free_port()), os.path.join(ROOT, 'bench.py'),
         '--impl', 'reference', '--gpus', '2', '--steps', '2', '--warmup',
         '1', '--mesh-n', '4'],
        capture_output=True, text=True, timeout=900, cwd=ROOT,
        env=dict(os.environ, OMP_NUM_THREADS='2')
    )
    assert res.returncode == 0, res.stderr[-2000:]
    _check(res.stdout, 2)
