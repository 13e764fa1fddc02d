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
        srv.settimeout(timeout)

        # Every other rank announces itself with its rank number; a stray
        # connection (or a rank connecting twice) is dropped instead of
        # consuming one of the size - 1 replies
        seen = set()
        try:
            while len(seen) < size - 1:
                conn, _ = srv.accept()
                with conn:
                    conn.settimeout(10)
                    try:
                        hello = conn.recv(16)
                        r = int(hello.decode().strip() or -1)
                    except (OSError, ValueError):
                        continue
                    if not (0 < r < size) or r in seen:
                        continue
                    conn.sendall(payload)
                    seen.add(r)
        except socket.timeout:
            raise TimeoutError(f'NCCL id rendezvous: only ranks '
                               f'{sorted(seen)} of {size - 1} reported '
                               f'within {timeout} s')
        finally:
            srv.close()

        return payload

    deadline = time.time() + timeout
    while True:
        try:
            with socket.create_connection((addr, port), timeout=5) as s:
                s.sendall(f'{rank:<16d}'.encode())
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


def _xchg_words(m):
    """Words moved for one exchange matrix: the dense ``[nrow][ncol]`` image.
    An ``align`` tag would pad the rows; such a matrix cannot be sent as one
    contiguous run."""
    if m.leaddim != m.ncol:
        raise ValueError('Exchange matrices must be tightly packed '
                         f'(leaddim {m.leaddim} != ncol {m.ncol})')
    return m.nrow*m.ncol


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
            n = _xchg_words(m)
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


class LoopbackWorld:
    """Several partitions in ONE process on ONE device.

    Every partition has its own backend (streams, graphs, exchange
    matrices); what would travel between ranks is copied device-to-device
    instead.  A send is a ``b200_memcpy_async`` of the packed ``XchgMatrix``
    into a mailbox buffer, issued on the stream the real exchange would use
    (and captured into the CUDA graph like it); a receive registers its
    destination.  ``deliver()`` plays the part of the transport: called
    between the stages of an RHS evaluation, after every partition has run
    its graph of that stage, it copies each mailbox into the matching
    receive buffer.  (It copies *every* registered pair each time; a
    receive buffer is only read in the stage that follows the one its data
    was sent in, so re-delivering older mailboxes is harmless.)

    This drives the partition-boundary path -- ``pack``, ``mpiconu``,
    ``mpicflux`` and the exchange matrices -- on a single GPU, with the
    semantics of ``pyfr/solvers/base/system.py:185-202`` and the requests
    of ``pyfr/backends/base/types.py:250-257``."""

    def __init__(self, size):
        self.size = size
        self.box, self.recvs = {}, {}
        self.rt = None

    def peer(self, rank):
        return LoopbackComm(self, rank)

    def deliver(self):
        rt = self.rt
        if rt is None:
            return

        rt.device_sync()
        for key, (dst, nb) in self.recvs.items():
            src = self.box.get(key)
            if src is not None:
                if src[1] != nb:
                    raise ValueError(f'Exchange {key}: sent {src[1]} bytes, '
                                     f'receiver expects {nb}')
                rt.memcpy(dst, src[0], nb)

    def run_lockstep(self, systems, t, uin, fout):
        """One RHS evaluation of all partitions, graph stage by stage."""
        for s in systems:
            for ks in s._get_kernels(uin, fout).values():
                for k in ks:
                    if k.rtnames:
                        k.bind(t=t)

        for stage in zip(*[s.rhs_graphs(uin, fout) for s in systems]):
            for s, g in zip(systems, stage):
                s.backend.run_graph(g)
            self.deliver()


class LoopbackComm:
    def __init__(self, world, rank):
        self.world, self.rank, self.size = world, rank, world.size
        self.rt = None

    def attach(self, rt):
        self.rt = self.world.rt = rt

    def register(self, r):
        """Called when a request is created (``XchgMatrix.sendreq`` /
        ``recvreq``): mailboxes are allocated here, outside any stream
        capture."""
        w, rt = self.world, self.rt
        m = r.mat
        nb = _xchg_words(m)*m.itemsize

        if r.kind == 'send':
            key = (self.rank, r.peer, r.tag)
            if key not in w.box:
                w.box[key] = (rt.new_ptr(rt.malloc, max(nb, 1)), nb)
        else:
            w.recvs[r.peer, self.rank, r.tag] = (m.data, nb)

    def exchange(self, reqs, stream):
        w, rt = self.world, self.rt

        for r in reqs:
            if r.kind == 'send':
                m = r.mat
                dst, nb = w.box[self.rank, r.peer, r.tag]
                rt.memcpy_async(dst, m.data, nb, stream)
