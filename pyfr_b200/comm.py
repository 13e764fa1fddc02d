"""Inter-partition communicator: NCCL point-to-point over NVLink.

Replaces the mpi4py persistent requests the reference attaches to every
``XchgMatrix`` (``pyfr/backends/base/types.py:250-257``; started and waited
on the host in ``pyfr/backends/cuda/types.py:99-116``).  One process per
GPU; rank 0 creates the NCCL unique id and hands it to the other ranks over
a one-shot TCP rendezvous on ``MASTER_ADDR:MASTER_PORT+17`` (the variables
``torchrun`` exports), so neither MPI nor torch is needed.  All sends and
receives of one RHS graph are issued as a single ``ncclGroupStart/End`` on
the backend's communication stream, directly on the device buffers.
"""

import ctypes as ct
import os
import socket
import sys
import time

import numpy as np

ID_BYTES = 128


def _rendezvous(rank, size, payload, addr, port, timeout=300):
    """Rank 0 serves ``payload`` to every other rank; returns it."""
    if rank == 0:
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((addr, port))
        srv.listen(size)

        for _ in range(size - 1):
            conn, _ = srv.accept()
            conn.sendall(payload)
            conn.close()

        srv.close()
        return payload

    deadline = time.time() + timeout
    while True:
        try:
            with socket.create_connection((addr, port), timeout=5) as s:
                buf = b''
                while len(buf) < ID_BYTES:
                    chunk = s.recv(ID_BYTES - len(buf))
                    if not chunk:
                        break
                    buf += chunk

                if len(buf) == ID_BYTES:
                    return buf
        except OSError:
            pass

        if time.time() > deadline:
            raise TimeoutError('NCCL id rendezvous timed out')

        time.sleep(0.2)


class NCCLComm:
    def __init__(self, rt, rank, size, addr=None, port=None):
        self.rt, self.rank, self.size = rt, rank, size

        addr = addr or os.environ.get('MASTER_ADDR', '127.0.0.1')
        port = int(port or os.environ.get('MASTER_PORT', 29500)) + 17

        uid = ct.create_string_buffer(ID_BYTES)
        if rank == 0:
            rt.nccl_unique_id(uid)

        raw = _rendezvous(rank, size, uid.raw, addr, port)

        # NCCL may print its version banner on stdout (NCCL_DEBUG set in
        # the environment); keep this process's stdout clean for callers
        # that parse it by pointing fd 1 at stderr while it initialises
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            self._handle = rt.new_ptr(rt.nccl_init, size, rank, raw)
        finally:
            os.dup2(saved, 1)
            os.close(saved)

    @classmethod
    def from_env(cls, rt):
        return cls(rt, int(os.environ.get('RANK', 0)),
                   int(os.environ.get('WORLD_SIZE', 1)))

    @staticmethod
    def _dtype(mat):
        return 1 if np.dtype(mat.dtype) == np.float64 else 0

    def exchange(self, reqs, stream):
        rt = self.rt
        rt.nccl_group_start()

        for r in reqs:
            m = r.mat
            n = m.nrow*m.ncol
            fn = rt.nccl_send if r.kind == 'send' else rt.nccl_recv
            fn(self._handle, m.data, n, self._dtype(m), r.peer, stream)

        rt.nccl_group_end()

    def allreduce(self, ptr, count, dtype_code, op, stream):
        """op: 0 = sum, 2 = max, 3 = min (ncclRedOp_t)."""
        self.rt.nccl_allreduce(self._handle, ptr, ptr, count, dtype_code, op,
                               stream)

    def close(self, destroy=False):
        """Release the communicator.

        ``ncclCommDestroy`` blocks for as long as CUDA graphs that captured
        operations on the communicator are alive (observed on NCCL 2.27:
        every rank hangs in it while the RHS graphs exist), and the RHS
        graphs live as long as the system object.  The default is therefore
        to leave the communicator to process exit; pass ``destroy=True``
        only after every graph that used it has been destroyed."""
        if self._handle and destroy:
            self.rt.nccl_destroy(self._handle)
        self._handle = None
