"""Generator for the multiply by a *dense* constant operator.

The operators of simplex-like elements (tetrahedra, pyramids; the
triangular factors of prisms) have no line structure: ``M0``, ``M4 -
M6*M0``, ``M6``, ``M1 - M3*M2`` and ``M3`` are 50-100 % dense, and ``out =
A @ b`` per element block is a small GEMM whose arithmetic (``M*K`` FMAs per
column) is on a par with its memory traffic.  The reference sends these to
cuBLASLt (``pyfr/backends/cuda/cublaslt.py:169-280``; GiMMiK declines them,
``gimmik.py:33-38``).  The sparse generator (``kernels/mul.py``) is the wrong
tool for them: a warp covers one 40-column row of a block with 24 of its 64
lanes idle, and every fp64 coefficient is a 64-bit literal costing two
uniform-register moves (profiles/r01z: 22 % of the HBM peak, 24 % of the
FP64 peak, register spills).

Here:

* a tile is ``NB`` consecutive element blocks, ``NB*LD`` a multiple of 32,
  so that every lane of every warp owns a column (fp64: 4 blocks = 160
  columns = 5 warps);
* a warp owns 32 columns and one group of output rows; per input row it
  reads one value from shared memory and issues one FMA per output row of
  its group, the coefficients coming from a copy of the operator in shared
  memory, read at warp-uniform addresses (one broadcast wavefront per two
  coefficients) -- 20-30 FMAs per data load -- in a *rolled* loop over the
  input rows, so the code is a few hundred instructions whatever the
  operator;
* the input tiles arrive by TMA bulk copy (one per block and chunk of input
  rows), double buffered on mbarriers, persistent CTAs.

Tensor cores are not used: B200's FP64 ``mma`` peak equals its FP64 FMA peak
(both ~40 TFLOP/s), so DMMA cannot beat an FMA kernel that keeps the FP64
pipe busy, and the 8x8x4 fragment shapes fit these operand shapes (K = 20-90,
40-column blocks) poorly.
"""

import numpy as np

from pyfr_b200.kernels import physics as ph
from pyfr_b200.kernels.mul import _pipeline_src


def is_dense(A, LD, isz, min_density=0.4, min_work=400):
    """Is ``A`` an operator this kernel should take?"""
    A = np.asarray(A)
    nnz = np.count_nonzero(A)
    return (nnz >= min_density*A.size and A.size >= min_work and
            _tile_blocks(LD) is not None and A.size <= 7600)


def _tile_blocks(LD):
    for nb in (1, 2, 4, 8):
        if (nb*LD) % 32 == 0:
            return nb
    return None


def dense_mul_source(be, A, LD, alpha, beta, negdiv_nvars=None,
                     rowgroups=3, max_tile_bytes=96*1024):
    """CUDA source for ``out = alpha*A@b + beta*out`` with a dense ``A``;
    optional ``negdivconf`` epilogue as in ``mul.mul_source``.  Returns
    (source, name, meta)."""
    A = alpha*np.asarray(A, dtype=float)
    M, K = A.shape
    isz = np.dtype(be.fpdtype).itemsize
    NB = _tile_blocks(LD)
    TC = NB*LD
    NWC = TC // 32

    # Row groups: enough warps to fill the SM, at most ~32 accumulators
    R = max(rowgroups, -(-M // 32))
    R = max(1, min(R, M, 32 // NWC))
    nthreads = 32*NWC*R

    # Chunks of input rows: two tiles of NB blocks within the budget.  A
    # block's tile is padded so that consecutive blocks start 16 banks
    # apart (a warp that straddles two blocks then touches 32 distinct
    # banks per half)
    maxrows_k = max(1, (max_tile_bytes // 2) // (TC*isz))
    nchunks = -(-K // maxrows_k)
    KC = -(-K // nchunks)
    chunks = [(k0, min(k0 + KC, K)) for k0 in range(0, K, KC)]
    wpr = LD*isz // 4                         # 32-bit words per row
    pad = 0 if (KC*wpr) % 32 == 16 else ((16 - (KC*wpr) % 32) % 32)*4 // isz
    BST = KC*LD + pad                         # block stride in the tile
    TILE = NB*BST

    # Coefficients in constant memory, transposed and padded: KA[k][m],
    # m padded to R equal row groups of RP rows, so that the RP coefficients
    # a warp needs for one input row are contiguous (two per 16-byte
    # uniform load) and the inner loops have fixed trip counts.  The loop
    # over the input rows stays *rolled*: fully unrolled, the thousands of
    # FMAs of a dense operator run once per tile and the kernel is bound by
    # instruction fetch (profiles/r02f: the unrolled form reached 22 % of
    # the FP64 peak, like the sparse generator it replaced).
    VW = 16 // isz                            # coefficients per load
    RP = -(-(-(-M // R)) // VW)*VW
    MP = R*RP
    vt = {2: 'double2', 4: 'float4'}[VW]
    lanes = 'xyzw'[:VW]
    AT = np.zeros((K, MP))
    AT[:, :M] = A.T
    cdecl = (f'static __device__ __align__(16) const fpdtype_t KA[{K*MP}] = {{'
             + ', '.join(ph.fpconst(v) for v in AT.ravel()) + '};')
    maxrows = RP

    def store(m, val):
        ix = f'ob + ({m})*LD'
        if negdiv_nvars:
            old = f'out[{ix}] + ' if beta == 1 else (
                f'{ph.fpconst(beta)}*out[{ix}] + ' if beta else '')
            return (f'out[{ix}] = -__ldg(rcpdjac + rjb + ({m})*C_SUB)*'
                    f'({old}{val});')
        if beta == 0:
            return f'out[{ix}] = {val};'
        elif beta == 1:
            return f'out[{ix}] += {val};'
        else:
            return f'out[{ix}] = fma({ph.fpconst(beta)}, out[{ix}], {val});'

    body = f'''
        const int kbeg = chunk*KC;
        const int kend = (kbeg + KC < {K}) ? kbeg + KC : {K};

        if (chunk == 0)
        {{
            UNROLL for (int i = 0; i < RP; i++)
                acc[i] = FP(0.0);
        }}

        #pragma unroll 2
        for (int k = kbeg; k < kend; k++)
        {{
            const fpdtype_t x = sm[(k - kbeg)*LD];
            const {vt} *ka = reinterpret_cast<const {vt} *>(
                KAs + k*{MP} + m0);
            UNROLL for (int i = 0; i < RP/{VW}; i++)
            {{
                const {vt} c = ka[i];
                {' '.join(f'acc[{VW}*i + {j}] = fma(c.{l}, x, acc[{VW}*i + {j}]);'
                          for j, l in enumerate(lanes))}
            }}
        }}

        if (chunk == NCHUNKS - 1 && live)
        {{
            UNROLL for (int i = 0; i < RP; i++)
                if (m0 + i < {M})
                {{
                    {store('m0 + i', 'acc[i]')}
                }}
        }}
'''

    extra_args = extra_pre = ''
    if negdiv_nvars:
        extra_args = (', const fpdtype_t* __restrict__ rcpdjac, '
                      'long long rcpdjac_bsz')
        extra_pre = (f'        const long long rjb = blk*rcpdjac_bsz + '
                     f'(cc/(K_SOA*{negdiv_nvars}))*K_SOA + cc % K_SOA;')

    # the operator itself, staged in shared memory once per CTA
    KAW = -(-K*MP*isz // 16)*16 // isz
    smem = (2*TILE + KAW)*isz + 16

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz)}
#define LD {LD}
#define NB {NB}
#define NCHUNKS {len(chunks)}
#define KC {KC}
#define BST {BST}
#define TILE {TILE}
#define NTHREADS {nthreads}
#define RP {RP}
{_pipeline_src}
{cdecl}

// out[{M} x LD] = A[{M} x {K}] @ b[{K} x LD] per element block, A dense
// ({int(np.count_nonzero(A))} non-zeros); tiles of {NB} blocks = {TC} columns
// = {NWC} warps x {R} row group(s), {len(chunks)} chunk(s) of {KC} input rows
extern "C" __global__ void __launch_bounds__(NTHREADS, 1)
opmul(int nblocks, const fpdtype_t* __restrict__ b, long long b_bsz,
      fpdtype_t* __restrict__ out, long long out_bsz{extra_args})
{{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    fpdtype_t *tiles = reinterpret_cast<fpdtype_t *>(smem_raw);
    fpdtype_t *KAs = tiles + 2*TILE;
    unsigned long long *full =
        reinterpret_cast<unsigned long long *>(KAs + {KAW});

    // The coefficients are read at warp-uniform addresses: from shared
    // memory that is one broadcast wavefront per 16 bytes.  (From constant
    // memory the {K*MP*isz//1024} KB operator cycles through the 2 KB first-level
    // constant cache once per tile: measured 15 % of the FP64 peak, r02j.)
    for (int i = threadIdx.x; i < {K*MP}; i += NTHREADS)
        KAs[i] = KA[i];

    const int tid = threadIdx.x;
    // (taken through a warp broadcast so that the compiler knows the row
    // group is uniform: the coefficient loads then go through the uniform
    // datapath, LDCU + FMA with a uniform-register operand, instead of one
    // per-thread constant load per FMA -- measured 3x slower, r02g)
    const int warp = __shfl_sync(0xffffffffu, tid / 32, 0);
    const int wc = warp % {NWC}, rg = warp / {NWC};
    const int m0 = rg*RP;

    // This thread's column: block jb of the tile, column cc of the block
    const int col = wc*32 + tid % 32;
    const int jb = col / LD, cc = col % LD;

    const long long ntiles = ((long long) nblocks + NB - 1) / NB;
    const long long mytiles = (ntiles - (long long) blockIdx.x
                               + gridDim.x - 1) / gridDim.x;
    const long long nitems = mytiles*NCHUNKS;

    if (tid == 0)
    {{
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }}
    __syncthreads();

    auto issue = [&](long long item)
    {{
        const long long b0 = (blockIdx.x + (item / NCHUNKS)*gridDim.x)*NB;
        const int chunk = (int) (item % NCHUNKS), st = (int) (item & 1);
        const int rows = (chunk == NCHUNKS - 1) ? {K} - chunk*KC : KC;
        const int nb = (int) ((nblocks - b0 < NB) ? nblocks - b0 : NB);

        mbar_expect_tx(&full[st], nb*rows*LD*sizeof(fpdtype_t));
        for (int j = 0; j < nb; j++)
            tma_load_1d(tiles + st*TILE + j*BST,
                        b + (b0 + j)*b_bsz + (long long) chunk*KC*LD,
                        rows*LD*sizeof(fpdtype_t), &full[st]);
    }};

    if (tid == 0 && nitems > 0)
        issue(0);

    fpdtype_t acc[{maxrows}];

    for (long long item = 0; item < nitems; item++)
    {{
        const int st = (int) (item & 1), chunk = (int) (item % NCHUNKS);
        const long long blk = (blockIdx.x + (item / NCHUNKS)*gridDim.x)*NB
                            + jb;
        const bool live = blk < nblocks;

        // Prefetch the next tile into the buffer released last iteration
        if (tid == 0 && item + 1 < nitems)
            issue(item + 1);

        mbar_wait(&full[st], (unsigned) ((item >> 1) & 1));

        const fpdtype_t *sm = tiles + st*TILE + jb*BST + cc;
        const long long ob = blk*out_bsz + cc;
{extra_pre}

{body}
        __syncthreads();
    }}
}}
'''
    meta = dict(nthreads=nthreads, nctas=1, smem=smem,
                nnz=int(np.count_nonzero(A)), nchunks=len(chunks), crows=KC,
                M=M, K=K, dense=True, nb=NB)

    return src, 'opmul', meta
