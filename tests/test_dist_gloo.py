"""CPU test of the N > 1 path: two real processes (torch.distributed,
gloo) each own one partition and exchange halos point to point."""

import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.parametrize('nranks', [2, 4])
def test_two_process_halo_exchange(nranks):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           f'--nproc-per-node={nranks}', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()),
           os.path.join(HERE, 'dist_oracle_worker.py')]
    env = dict(os.environ, OMP_NUM_THREADS='1')
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600,
                         env=env)

    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-2000:])
    assert res.stdout.count('bit-identical=True') == 3*nranks
    assert res.stdout.count('adaptive: decisions=True dt=True') == nranks
