"""Device-resident matrices, views, exchange buffers and the run graph.

Concrete B200 versions of the ``pyfr.backends.base`` types (compare the
reference's CUDA backend, ``pyfr/backends/cuda/types.py:14-116``).
Differences that matter:

* exchange matrices live on the device only; the inter-partition exchange
  is a grouped NCCL send/recv on a communication stream over NVLink, so
  there are no pinned bounce buffers and no host ``Waitall``;
* a graph is recorded once by stream capture (kernels on the compute
  stream, the exchange forked onto the communication stream after the last
  pack kernel and joined at the end) and replayed as a single CUDA-graph
  launch.
"""

import numpy as np

from pyfr_b200 import base


class DevAlloc:
    """Owns a device allocation; freed when the last reference dies."""

    def __init__(self, rt, nbytes):
        self.rt, self.nbytes = rt, nbytes
        self.ptr = rt.new_ptr(rt.malloc, nbytes)
        if not rt.dry:
            # The fill runs on the legacy stream, which the backend's
            # non-blocking streams do not synchronise with: wait for it
            rt.memset(self.ptr, 0, nbytes, None)
            rt.stream_sync(None)

    def __int__(self):
        return self.ptr

    def __del__(self):
        try:
            self.rt.free(self.ptr)
        except Exception:
            pass


class _Common:
    @property
    def data(self):
        return int(self.basedata) + self.offset

    @property
    def _as_parameter_(self):
        return self.data


class B200MatrixBase(_Common, base.MatrixBase):
    def onalloc(self, basedata, offset):
        self.basedata, self.offset = basedata, offset

        if self._initval is not None:
            self._hostcopy = self._initval if 'const' in self.tags else None
            self._set(self._initval)

        del self._initval

    def _get(self):
        if getattr(self, '_hostcopy', None) is not None:
            return self._hostcopy

        rt = self.backend.rt
        buf = np.empty(self.nbytes // self.itemsize, dtype=self.dtype)
        rt.device_sync()
        rt.memcpy(buf.ctypes.data, self.data, self.nbytes)

        return self._unpack(buf)

    def _set(self, ary):
        rt = self.backend.rt
        if rt.dry:
            return

        buf = np.ascontiguousarray(self._pack(ary))
        rt.device_sync()
        rt.memcpy(self.data, buf.ctypes.data, self.nbytes)

    # Raw transfers of the storage-order image (what the reference moves
    # through its pinned bounce buffer, cuda/types.py:25-41), asynchronous
    # on the backend's stream; ``hostptr`` should be pinned memory
    def upload_packed(self, hostptr, stream=None):
        be = self.backend
        be.rt.memcpy_async(self.data, hostptr, self.nbytes,
                           stream if stream is not None else be.stream)

    def download_packed(self, hostptr, stream=None):
        be = self.backend
        be.rt.memcpy_async(hostptr, self.data, self.nbytes,
                           stream if stream is not None else be.stream)


class B200Matrix(B200MatrixBase, base.Matrix): pass


class B200ConstMatrix(B200MatrixBase, base.ConstMatrix):
    def __init__(self, backend, dtype, initval, tags):
        # Unaligned 2-D constant tables (normals, index maps, point sets)
        # are only ever addressed point-wise, so they are kept row-major
        if initval.ndim == 2 and 'align' not in tags:
            tags = set(tags) | {'noblock'}

        base.ConstMatrix.__init__(self, backend, dtype, initval, tags)


class B200MatrixSlice(_Common, base.MatrixSlice): pass
class B200View(base.View): pass
class B200XchgView(base.XchgView): pass


class XchgRequest:
    """One direction of one neighbour's halo exchange."""

    def __init__(self, kind, mat, peer, tag):
        self.kind, self.mat, self.peer, self.tag = kind, mat, peer, tag


class B200XchgMatrix(B200Matrix, base.XchgMatrix):
    def recvreq(self, comm, pid, tag):
        return self._req('recv', comm, pid, tag)

    def sendreq(self, comm, pid, tag):
        return self._req('send', comm, pid, tag)

    def _req(self, kind, comm, pid, tag):
        req = XchgRequest(kind, self, pid, tag)
        # A communicator may want to know its requests before the first
        # exchange runs (allocations are not allowed during stream capture)
        if hasattr(comm, 'register'):
            comm.register(req)
        return req


class B200Graph(base.Graph):
    """Kernels + exchanges in schedule order, replayed via CUDA graphs."""

    def __init__(self, backend):
        super().__init__(backend)
        self.program = []
        self._exec = None

    def _add_mpi_req(self, req, deps):
        super()._add_mpi_req(req, deps)
        self.program.append(('xchg', req))

    def note_kernel(self, kern):
        self.program.append(('kernel', kern))
        return kern

    def _group(self, kerns, subs):
        self._fgroups = getattr(self, '_fgroups', []) + [(kerns, subs)]

    def _fuse(self):
        be = self.backend
        if not be.fuse:
            return

        # (the graph being planned, for rewrites that look at its other
        # kernels: fusion.conu_fold_plan)
        be._fusing = self
        try:
            self._fuse_groups(be)
        finally:
            be._fusing = None

    def _fuse_groups(self, be):
        from pyfr_b200 import fusion

        for kerns, subs in getattr(self, '_fgroups', []):
            new = fusion.fuse_group(be, kerns, subs)
            if not new and kerns and \
               getattr(kerns[-1], 'kind', None) == 'rkvdh2':
                # the RHS chain fuses without the stage update
                kerns = kerns[:-1]
                new = fusion.fuse_group(be, kerns, subs)
            if not new:
                continue

            old = {id(l) for k in kerns for l in fusion.leaves(k)}
            prog, done = [], False
            for what, obj in self.program:
                if what == 'kernel' and id(obj) in old:
                    if not done:
                        prog.extend(('kernel', k) for k in new)
                        done = True
                else:
                    prog.append((what, obj))

            self.program = prog

        self.program = fusion.elide_copy_fpts(be, self.program)

        if be.batch_launches:
            self.program = fusion.batch_launches(be, self.program)

    def _commit(self):
        import weakref

        self._fuse()
        self._plan()

        # (the fusion pass of the next graph may still fold a kernel of
        # this one into its own -- ``fusion.conu_fold_plan`` -- as long as
        # this graph has not run)
        self.backend.last_committed = weakref.ref(self)

    def _plan(self):
        # All exchanges of a graph go out as one NCCL group, issued once
        # the last pack kernel has been enqueued
        reqs = [r for what, r in self.program if what == 'xchg']
        kerns = [k for what, k in self.program if what == 'kernel']

        plan = []
        if reqs:
            last = max((i for i, (w, r) in enumerate(self.program)
                        if w == 'xchg'), default=-1)

            # A graph that only receives has no pack kernel to anchor the
            # exchange to and the base scheduler lists its receives first.
            # Posting them there would park the NCCL kernel's CTAs on some
            # SMs, spinning until the peer sends (which it does only after
            # *its* element kernel), and the persistent element kernel of
            # this rank -- one CTA per SM, all of the shared memory -- could
            # not be co-resident with them: it would run in two waves.  So
            # a receive-only exchange is issued where the sending side
            # issues its half: just before the last kernel of the graph.
            if all(r.kind == 'recv' for r in reqs) and kerns:
                last = max(i for i, (w, k) in enumerate(self.program)
                           if w == 'kernel' and k is kerns[-1]) - 1
                while last >= 0 and self.program[last][0] != 'kernel':
                    last -= 1

            issued = False
            for i, (what, obj) in enumerate(self.program):
                if last < 0 and not issued:
                    plan.append(('xchg', reqs))
                    issued = True
                if what == 'kernel':
                    plan.append(('kernel', obj))
                if i == last and not issued:
                    plan.append(('xchg', reqs))
                    issued = True
        else:
            plan = [('kernel', k) for k in kerns]

        self.plan = plan

    def _record(self, stream):
        be, rt = self.backend, self.backend.rt
        joined = True

        for what, obj in self.plan:
            if what == 'kernel':
                obj.run(stream)
            else:
                # Fork: communication stream waits for the packs
                rt.event_record(be.fork_event, stream)
                rt.stream_wait_event(be.comm_stream, be.fork_event)
                be.exchange(obj, be.comm_stream)
                rt.event_record(be.join_event, be.comm_stream)
                joined = False

        if not joined:
            rt.stream_wait_event(stream, be.join_event)

    def run(self, stream=None):
        be, rt = self.backend, self.backend.rt
        stream = stream or be.stream
        self.started = True

        # Run-time scalars (t, dt) reach the kernels through device memory
        be.rtscal.flush(stream)

        if not be.use_graphs:
            self._record(stream)
            return

        if self._exec is None or any(getattr(k, 'dirty', False)
                                     for w, k in self.plan if w == 'kernel'):
            if self._exec is not None:
                rt.graph_destroy(self._exec)

            rt.capture_begin(stream)
            self._record(stream)
            self._exec = rt.end_capture(stream)

            for w, k in self.plan:
                if w == 'kernel':
                    k.dirty = False

        rt.graph_launch(self._exec, stream)

    def __del__(self):
        if getattr(self, '_exec', None) is not None:
            try:
                self.backend.rt.graph_destroy(self._exec)
            except Exception:
                pass
