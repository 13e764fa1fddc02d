"""The ``b200`` backend: PyFR's backend contract on one NVIDIA B200.

Plays the role of ``CUDABackend`` (``pyfr/backends/cuda/base.py:9-144``)
with a different design underneath: blocked AoSoA storage
(``blocks = True``, so one element block is a contiguous tile that a single
TMA bulk copy moves), operator constants baked into generated sm_100a
kernels, CUDA-graph replay of each RHS graph and NCCL for the halo
exchange.  Configuration keys (section ``[backend-b200]``): ``device-id``
(``local-rank`` or an index), ``n-soa`` (SoA width, default 64 bytes worth
of scalars), ``n-csub`` (columns per block, multiple of ``n-soa``),
``graphs`` (replay through CUDA graphs, default on), ``fusion`` (replace
grouped kernel chains by fused single-launch kernels, default on).
"""

import os

import numpy as np

from pyfr_b200 import base, providers, types
from pyfr_b200.compiler import KernelCompiler
from pyfr_b200.lib import load_runtime


class RuntimeScalars:
    """Run-time scalar kernel arguments (``t`` of time-dependent boundary
    data, ``dt`` of a stage update) kept in device memory.

    The reference updates such arguments on the executable graph's nodes
    (``pyfr/backends/cuda/types.py:86-97``, ``cuGraphExecKernelNodeSetParams``).
    Here a kernel takes a *pointer* to its scalar instead: ``bind`` writes
    the host mirror, ``flush`` (issued before a graph is launched) uploads
    the changed range with one small copy on the compute stream, and the
    captured graph is replayed unchanged -- no re-capture when ``t`` or
    ``dt`` moves.  The host mirror is pageable memory on purpose: the copy
    is then staged at call time, so the next ``bind`` cannot race with a
    copy that has not executed yet."""

    nslots = 4096

    def __init__(self, be):
        self.be = be
        self.host = np.zeros(self.nslots, dtype=be.fpdtype)
        self.dev = None
        self.nused = 0
        self.lo, self.hi = self.nslots, 0

    def alloc(self):
        if self.nused == self.nslots:
            raise RuntimeError('Out of run-time scalar slots')
        if self.dev is None:
            self.dev = types.DevAlloc(self.be.rt, self.host.nbytes)
        self.nused += 1
        return self.nused - 1

    def ptr(self, idx):
        return int(self.dev) + idx*self.host.itemsize

    def set(self, idx, value):
        # compared in the storage type (a float32 slot holding 0.1f does
        # not differ from the Python double 0.1 bound again)
        v = self.host.dtype.type(value)
        if self.host[idx] != v:
            self.host[idx] = v
            self.lo, self.hi = min(self.lo, idx), max(self.hi, idx + 1)

    def flush(self, stream):
        if self.hi > self.lo and not self.be.rt.dry:
            isz = self.host.itemsize
            self.be.rt.memcpy_async(int(self.dev) + self.lo*isz,
                                    self.host.ctypes.data + self.lo*isz,
                                    (self.hi - self.lo)*isz, stream)
        self.lo, self.hi = self.nslots, 0


class B200Backend(base.BaseBackend):
    name = 'b200'
    blocks = True

    const_matrix_cls = types.B200ConstMatrix
    matrix_cls = types.B200Matrix
    matrix_slice_cls = types.B200MatrixSlice
    view_cls = types.B200View
    xchg_matrix_cls = types.B200XchgMatrix
    xchg_view_cls = types.B200XchgView
    graph_cls = types.B200Graph
    ordered_meta_kernel_cls = providers.B200OrderedMetaKernel
    unordered_meta_kernel_cls = providers.B200UnorderedMetaKernel

    def __init__(self, cfg, dry=False, comm=None):
        super().__init__(cfg)
        sect = 'backend-b200'

        devid = cfg.get(sect, 'device-id', 'local-rank')
        if devid == 'local-rank':
            devid = int(os.environ.get('LOCAL_RANK', 0))

        self.rt = rt = load_runtime(int(devid), dry=dry)
        info = rt.device_info()

        if not dry and info['cc'][0] != 10:
            raise RuntimeError('The b200 backend targets sm_100a only; found '
                               f'compute capability {info["cc"]}')

        # (persistent kernels launch one CTA, or CTA pair, per SM; `sm-count`
        # restricts them to fewer -- a test and tuning knob)
        self.sm_count = cfg.getint(sect, 'sm-count', info['sm_count'])
        self.smem_budget = min(info['smem_optin'], 227*1024) - 8*1024

        # Storage layout: 64-byte SoA rows (two 32-byte sectors), one SoA
        # group per block; small enough that the fused element kernel can
        # hold a whole block's gradients in shared memory at p = 4
        isz = np.dtype(self.fpdtype).itemsize
        self.alignb = 256
        if cfg.hasopt(sect, 'n-soa'):
            self.soasz = cfg.getint(sect, 'n-soa')
        else:
            self.soasz = self._auto_soasz(cfg, isz)
        self.csubsz = cfg.getint(sect, 'n-csub', self.soasz)
        if self.csubsz % self.soasz:
            raise ValueError('n-csub must be a multiple of n-soa')

        # (options of the table-driven fused kernel select that kernel;
        # looked up before the getters below record their defaults)
        table_opts = any(cfg.hasopt(sect, o) for o in (
            'gradflux-vec2', 'gradflux-planes', 'gradflux-ncol',
            'gradflux-monojac'))

        # row groups (= threads) of the sparse operator kernel; 0: four for a
        # pure stream (out = A b), eight where the kernel also reads `out`
        # (beta != 0, negdivconf / stage-update epilogues: more loads in
        # flight -- 0.796 -> 0.641 ms for M3 + negdivconf, while M0 loses
        # 10 % at eight; r02k)
        self.mul_rowgroups = cfg.getint(sect, 'mul-rowgroups', 0)
        # dense operators (tets, pyramids) take the small-GEMM kernel
        self.dense_mul = cfg.getbool(sect, 'dense-mul', True)
        # ... on the FP64 tensor cores (mma.sync m8n8k4) in double precision
        self.dense_mma = cfg.getbool(sect, 'dense-mma', True)
        # fp64 operators with at least this many distinct coefficients keep
        # them in __constant__ memory (0: always literals)
        self.mul_const_table = cfg.getint(sect, 'mul-const-table', 0)
        self.cflux_minblocks = cfg.getint(sect, 'cflux-minblocks', 5)
        self.gradflux_maxctas = cfg.getint(sect, 'gradflux-maxctas', 2)
        self.gradflux_threads = cfg.getint(sect, 'gradflux-threads', 0)
        # sum-factorised fused kernel for tensor-product elements
        self.gradflux_tensor = (cfg.getbool(sect, 'gradflux-tensor', True)
                                and not table_opts)
        # ... its warp groups (each owns half of a block's elements) and the
        # phase after which the second group starts (1 or 3)
        self.gradflux_groups = cfg.getint(sect, 'gradflux-groups', 1)
        self.gradflux_stagger = cfg.getint(sect, 'gradflux-stagger', 1)
        # loads of the next work item written ahead of the arithmetic of
        # the current one in the line phases (tensor-product kernel)
        self.gradflux_swp = cfg.getbool(sect, 'gradflux-swp', True)
        # fetch the constant metric of a flux-point work item after its
        # interpolation (fewer live registers, three fetches per block)
        self.gradflux_metric_late = cfg.getbool(sect, 'gradflux-metric-late',
                                                False)
        self.gradflux_planes = cfg.getbool(sect, 'gradflux-planes', False)
        self.gradflux_monojac = cfg.getbool(sect, 'gradflux-monojac', True)
        self.gradflux_ncol = cfg.getint(sect, 'gradflux-ncol', 1)
        # Phases of gradflux whose work items take two adjacent columns
        # with 16-byte accesses ('p1', 'p3', 'p5'; comma separated)
        self.gradflux_vec2 = tuple(
            v.strip() for v in str(cfg.get(sect, 'gradflux-vec2', '')).split(',')
            if v.strip() and v.strip() != '0'
        )
        # intconu: two consecutive interface points per thread, 128-bit
        # accesses where a side's addresses are adjacent and aligned
        self.conu_pairs = cfg.getbool(sect, 'conu-pairs', False)
        # order of the points of interior / boundary interfaces chosen by
        # the host mirror: 'reference' (the reference's sort key) or
        # 'address' (true left-hand address)
        self.inters_order = cfg.get(sect, 'inters-order', 'reference')
        if self.inters_order not in ('reference', 'address'):
            raise ValueError('inters-order must be reference or address')
        # order in which the interior / boundary interface kernels visit
        # their points: 'address' (sorted by left-hand address inside the
        # provider; the views the host built are left untouched) or 'host'
        self.kernel_order = cfg.get(sect, 'kernel-order', 'address')
        if self.kernel_order not in ('address', 'host'):
            raise ValueError('kernel-order must be address or host')
        self.affine_fastpath = cfg.getbool(sect, 'affine-fastpath', True)
        self.euler_fusion = cfg.getbool(sect, 'euler-fusion', True)
        # Runge-Kutta stage update in the epilogue of the last RHS kernel
        # (only when the caller groups an rkvdh2 kernel with the RHS)
        self.rk_fusion = cfg.getbool(sect, 'rk-fusion', True)
        # one launch for the per-neighbour kernels (pack, mpiconu, mpicflux)
        self.batch_launches = cfg.getbool(sect, 'batch-launches', True)
        self.use_graphs = cfg.getbool(sect, 'graphs', True) and not dry
        self.fuse = cfg.getbool(sect, 'fusion', True)

        self.compiler = KernelCompiler(rt)
        self.nlaunches = 0
        self.view_uses = []
        self._ordered = {}
        self.dead_rows = cfg.getbool(sect, 'dead-rows', True)
        # interior common solution (|ldg-beta| = 1/2) gathered by the
        # element kernel: no intconu launch (fusion.conu_fold_plan)
        self.conu_fold = cfg.getbool(sect, 'conu-fold', True)
        # ... whole flux-point rows by bulk copy where their points come
        # from one row of the trace matrix
        self.gather_rows = cfg.getbool(sect, 'gather-rows', True)
        # element kernel on half blocks, two CTAs per SM, where a whole
        # block fills the shared memory (tensor-product kernel)
        # (measured a loss with per-thread copies in place of the bulk
        # copy: 0.275 -> 0.281 ms at 32^3, r02n; off)
        self.gradflux_split = cfg.getbool(sect, 'gradflux-split', False)
        self.last_committed = None
        # partitioned meshes: the element kernel split into boundary and
        # interior blocks, the latter behind the exchange of the traces
        self.gradflux_overlap = cfg.getbool(sect, 'gradflux-overlap', True)

        # Compute stream, communication stream and fork/join events
        self.stream = rt.new_ptr(rt.stream_create)
        self.comm_stream = rt.new_ptr(rt.stream_create_priority, 1)
        self.fork_event = rt.new_ptr(rt.event_create)
        self.join_event = rt.new_ptr(rt.event_create)

        self.rtscal = RuntimeScalars(self)

        self.comm = comm
        if hasattr(comm, 'attach'):
            comm.attach(rt)
        self.pointwise = providers.PointwiseProvider(self)
        self._providers = [providers.OperatorProvider(self),
                           providers.BlasExtProvider(self),
                           providers.PackingProvider(self), self.pointwise]

    @staticmethod
    def _auto_soasz(cfg, isz, smem=212*1024):
        """SoA width: 64 bytes' worth of scalars, halved (down to one
        16-byte access) until the fused element kernel of a hexahedron of
        the configured order keeps a whole block -- solution, common
        solution and the ndims gradient components of nvars = 5 fields --
        in shared memory.  fp64 p <= 4 and fp32 p <= 3 keep the full width;
        fp32 p = 6 (BASELINE configs[4]) runs at a width of 4, where the
        fused kernel is 2.2x faster than the eleven-launch chain of the
        full width (profiles/r02e)."""
        soasz = 64 // isz
        try:
            n1 = cfg.getint('solver', 'order') + 1
        except Exception:
            return soasz

        rows = 4*n1**3 + 6*n1**2
        while soasz*isz > 16 and rows*5*soasz*isz > smem:
            soasz //= 2
        return soasz

    def _malloc_impl(self, nbytes):
        return types.DevAlloc(self.rt, nbytes)

    def run_kernels(self, kernels, wait=False):
        self.rtscal.flush(self.stream)
        for k in kernels:
            k.run(self.stream)

        if wait:
            self.wait()

    def run_graph(self, graph, wait=False):
        graph.run(self.stream)

        if wait:
            self.wait()

    def wait(self):
        self.rt.stream_sync(self.stream)

    def exchange(self, reqs, stream):
        """Issue all sends/receives of one graph as a single NCCL group."""
        if self.comm is None:
            raise RuntimeError('Inter-partition exchange requested but the '
                               'backend has no communicator')

        self.comm.exchange(reqs, stream)

    def memory_info(self):
        info = self.rt.device_info()
        cur = super().memory_info()
        return base.MemoryInfo(cur.current, cur.peak, info['free_mem'],
                               info['total_mem'])
