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

Measured (profiles/r02k, mixed hex+pri+pyr+tet mesh, 24^3 cells, p = 3): the
dense products run at 26-30 % of the FP64 peak (``mul[56x30]`` 0.127 -> 0.062 ms,
``mul[90x30]`` 0.22 -> 0.086 ms against the sparse generator), the whole RHS
4.63 -> 6.22 GDoF/s.  What bounds them now is *operand delivery*, found by
elimination: coefficients as per-thread constant loads (``LDC``) or through
the uniform datapath (``LDCU``) cycle a 13-40 KB operator through the 2 KB
first-level constant cache once per tile (15 %); fully unrolled with immediate
constant-bank operands the kernel is bound by instruction fetch (22 %); from
shared memory every 16-byte broadcast still writes 512 bytes of registers per
warp (4 cycles) for two FMAs' worth of coefficients.  An FMA formulation
needs one operand delivered per FMA; ``mma`` does not -- a DMMA 8x8x4 performs
256 FMAs on one 8-byte fragment load per lane and operand.  Tensor cores
would not add FLOPs here (B200's FP64 ``mma`` peak equals its FMA peak) but
they would lift this operand bound: that is the next step for these shapes.
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


def dense_mma_source(be, A, LD, alpha, beta, negdiv_nvars=None,
                     max_tile_bytes=96*1024):
    """The dense multiply on the FP64 tensor cores (``mma.sync.m8n8k4``).

    Same tiles and TMA pipeline as ``dense_mul_source``; a warp owns a
    ``MW x 32`` patch of the output tile (``MW`` = 16 or 32 operator rows,
    32 columns) as ``MW/8 x 4`` accumulator fragments and, per step of four
    input rows, loads ``MW/8`` operator fragments and four data fragments
    from shared memory -- one 8-byte value per lane each -- for ``MW/2``
    ``mma`` instructions of 256 FMAs.  The operator is staged in shared
    memory once per CTA, rows padded to a stride of 4 (mod 16) doubles so
    that a fragment load is conflict free.  fp64 only."""
    A = alpha*np.asarray(A, dtype=float)
    M, K = A.shape
    isz = np.dtype(be.fpdtype).itemsize
    if isz != 8:
        raise ValueError('the mma form is fp64 only')

    NB = _tile_blocks(LD)
    TC = NB*LD
    if LD % 8 or TC % 32:
        raise ValueError('columns do not tile into 8-wide fragments')
    NWC = TC // 32

    MW = 32 if M > 48 else 16
    R = -(-M // MW)
    FM, FN = MW // 8, 4
    nthreads = 32*NWC*R

    # input rows in chunks (multiples of four rows), two tiles in flight
    maxrows_k = max(4, ((max_tile_bytes // 2) // (TC*isz)) // 4 * 4)
    nchunks = -(-K // maxrows_k)
    KC = -(-(-(-K // nchunks)) // 4)*4
    nchunks = -(-K // KC)
    BST = KC*LD + 8                            # (+8: see dense_mul_source)
    if (KC*LD*2) % 32 == 16:
        BST = KC*LD
    TILE = NB*BST

    K4 = -(-K // 4)*4
    KP = K4 + (4 - K4 % 16) % 16               # row stride = 4 (mod 16)
    MP = R*MW
    AP = np.zeros((MP, KP))
    AP[:M, :K] = A
    cdecl = (f'static __device__ __align__(16) const fpdtype_t '
             f'KA[{MP*KP}] = {{'
             + ', '.join(ph.fpconst(v) for v in AP.ravel()) + '};')

    def store(val, m, c):
        ix = f'ob + ({m})*LD + {c}'
        if negdiv_nvars:
            old = f'out[{ix}] + ' if beta == 1 else (
                f'{ph.fpconst(beta)}*out[{ix}] + ' if beta else '')
            return (f'out[{ix}] = -__ldg(rcpdjac + rjb + ({m})*C_SUB + '
                    f'RJO({c}))*({old}{val});')
        if beta == 0:
            return f'out[{ix}] = {val};'
        elif beta == 1:
            return f'out[{ix}] += {val};'
        else:
            return f'out[{ix}] = fma({ph.fpconst(beta)}, out[{ix}], {val});'

    extra_args = ''
    if negdiv_nvars:
        extra_args = (', const fpdtype_t* __restrict__ rcpdjac, '
                      'long long rcpdjac_bsz')

    smem = (2*TILE + MP*KP)*isz + 16
    # two CTAs per SM where the tiles are small (few row groups)
    nctas = 2 if (2*(smem + 1024) <= 227*1024 and 2*nthreads <= 768) else 1

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz)}
#define LD {LD}
#define NB {NB}
#define NCHUNKS {nchunks}
#define KC {KC}
#define BST {BST}
#define TILE {TILE}
#define NTHREADS {nthreads}
#define FM {FM}
#define FN {FN}
#define KP {KP}
#define RJO(c) (((c)/(K_SOA*{negdiv_nvars or 1}))*K_SOA + (c) % K_SOA)
{_pipeline_src}
{cdecl}

// out[{M} x LD] = A[{M} x {K}] @ b[{K} x LD] per element block on the FP64
// tensor cores; tiles of {NB} blocks = {TC} columns = {NWC} warps x {R} row
// group(s) of {MW} rows, {nchunks} chunk(s) of {KC} input rows
extern "C" __global__ void __launch_bounds__(NTHREADS, {nctas})
opmul(int nblocks, const fpdtype_t* __restrict__ b, long long b_bsz,
      fpdtype_t* __restrict__ out, long long out_bsz{extra_args})
{{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    fpdtype_t *tiles = reinterpret_cast<fpdtype_t *>(smem_raw);
    fpdtype_t *KAs = tiles + 2*TILE;
    unsigned long long *full =
        reinterpret_cast<unsigned long long *>(KAs + {MP*KP});

    for (int i = threadIdx.x; i < {MP*KP}; i += NTHREADS)
        KAs[i] = KA[i];

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid / 32, 0);
    const int wc = warp % {NWC}, rg = warp / {NWC};
    const int lane = tid % 32, g = lane / 4, t = lane % 4;
    const int m0 = rg*{MW}, n0 = wc*32;

    // operator fragments: A[m0 + 8i + g][4s + t]
    const fpdtype_t *ap = KAs + (m0 + g)*KP + t;

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

    // data fragments: column n0 + 8j + g of the tile = column cc of block jb
    int boff[FN];
    UNROLL for (int j = 0; j < FN; j++)
    {{
        const int n = n0 + 8*j + g;
        boff[j] = (n / LD)*BST + n % LD + t*LD;
    }}

    fpdtype_t acc[FM][FN][2];

    for (long long item = 0; item < nitems; item++)
    {{
        const int st = (int) (item & 1), chunk = (int) (item % NCHUNKS);
        const long long blk0 = (blockIdx.x + (item / NCHUNKS)*gridDim.x)*NB;

        if (tid == 0 && item + 1 < nitems)
            issue(item + 1);

        mbar_wait(&full[st], (unsigned) ((item >> 1) & 1));

        const fpdtype_t *sm = tiles + st*TILE;
        const int kbeg = chunk*KC;
        const int krows = ((kbeg + KC < {K}) ? KC : {K} - kbeg);

        if (chunk == 0)
        {{
            UNROLL for (int i = 0; i < FM; i++)
                UNROLL for (int j = 0; j < FN; j++)
                    acc[i][j][0] = acc[i][j][1] = FP(0.0);
        }}

        #pragma unroll 2
        for (int s = 0; s < (krows + 3)/4; s++)
        {{
            fpdtype_t af[FM], bf[FN];
            UNROLL for (int i = 0; i < FM; i++)
                af[i] = ap[8*i*KP + kbeg + 4*s];
            // (rows past the operator's last column: the padded
            // coefficient is zero, but the tile holds nothing there)
            const bool in = 4*s + t < krows;
            UNROLL for (int j = 0; j < FN; j++)
                bf[j] = in ? sm[boff[j] + 4*s*LD] : FP(0.0);

            UNROLL for (int i = 0; i < FM; i++)
                UNROLL for (int j = 0; j < FN; j++)
                    mma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }}

        if (chunk == NCHUNKS - 1)
        {{
            // this lane's entries: rows m0 + 8i + g, columns n0 + 8j + 2t, +1
            UNROLL for (int j = 0; j < FN; j++)
            {{
                const int n = n0 + 8*j + 2*t;
                const long long blk = blk0 + n / LD;
                const int cc = n % LD;
                if (blk < nblocks)
                {{
                    const long long ob = blk*out_bsz;
                    {'const long long rjb = blk*rcpdjac_bsz;'
                     if negdiv_nvars else ''}
                    UNROLL for (int i = 0; i < FM; i++)
                    {{
                        const int m = m0 + 8*i + g;
                        if (m < {M})
                        {{
                            {store('acc[i][j][0]', 'm', 'cc')}
                            {store('acc[i][j][1]', 'm', 'cc + 1')}
                        }}
                    }}
                }}
            }}
        }}

        __syncthreads();
    }}
}}
'''
    meta = dict(nthreads=nthreads, nctas=nctas, smem=smem,
                nnz=int(np.count_nonzero(A)), nchunks=nchunks, crows=KC,
                M=M, K=K, dense=True, mma=True, nb=NB)

    return src, 'opmul', meta
