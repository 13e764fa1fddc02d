"""Generator for the fused Navier-Stokes element kernel of tensor-product
elements (hexahedra, quadrilaterals): ``gradflux`` in sum-factorised form.

Same contract as ``fused.gradflux_source`` -- one launch for the chain
``tgradpcoru_upts .. tdivtpcorf_upts`` of ``pyfr/solvers/baseadvecdiff/
system.py:94-205`` -- but written for what the operators of a
tensor-product element *are* (``pyfr/shapes.py:79-135`` builds them as
Kronecker products of 1-D matrices): every operator acts along the lines
of solution points of one direction, with one small 1-D matrix per
direction.  The generator verifies that structure on the matrices it is
given (it rebuilds them from the extracted 1-D factors and compares) and
declines otherwise, so nothing here depends on how the host numbered the
points.

What this buys over the table-driven kernel (ncu, profiles/r01u: 516 M
shared-memory wavefronts, 1.31 G warp instructions, 58 instructions per
flux-point output):

* a thread owns a fixed set of (line, column-group) work items for the
  whole launch, so every row offset is formed once per kernel and lives in
  a register: no index tables, no per-block address arithmetic;
* a work item covers ``16/sizeof(fp)`` adjacent columns with one 16-byte
  shared-memory access per row (``LDS.128``/``STS.128``/``STG.128``);
* for affine elements (constant Jacobian) the gradient is never
  transformed in shared memory: interpolation to the flux points and
  multiplication by the constant metric commute, so the flux-point pass
  combines the three interpolated reference gradients in registers and the
  flux pass transforms its own point -- the separate ``gradcoru`` pass over
  ``G`` disappears;
* the last direction of the divergence is fused with the final sum, which
  streams straight to HBM.

Per element block (hex, p = 4, fp64) the kernel moves 3650 shared-memory
words per column instead of ~6300.
"""

import numpy as np

from pyfr_b200.kernels import physics as ph
from pyfr_b200.kernels.fused import ConstPool, NotFusable, geometry_source
from pyfr_b200.kernels.mul import _pipeline_src


def _components(adj):
    n = len(adj)
    seen, comps = np.zeros(n, dtype=bool), []

    for p0 in range(n):
        if seen[p0]:
            continue

        comp, todo = [], [p0]
        seen[p0] = True
        while todo:
            p = todo.pop()
            comp.append(p)
            for q in np.flatnonzero(adj[p]):
                if not seen[q]:
                    seen[q] = True
                    todo.append(q)

        comps.append(sorted(comp))

    return comps


def tp_structure(ops, nd, tol=1e-12):
    """Line structure of the operators of a tensor-product element.

    ``ops``: ``A1`` (``M4 - M6*M0``, ndims*nupts x nupts), ``M6``
    (ndims*nupts x nfpts), ``M0`` (nfpts x nupts), ``A5`` (``M1 - M3*M2``,
    nupts x ndims*nupts).  Returns a dict with, per direction ``d``: the
    lines (first row and row stride of their ``n1`` solution points, the
    flux-point rows at their two ends) and the 1-D matrices ``Dg`` (n1 x
    n1), ``Lg`` (n1 x 2), ``lm``/``lp`` (n1) and ``Dt`` (n1 x n1) such that

        (A1 u + M6 c)[d, line] = Dg[d] u[line] + Lg[d] (c[f-], c[f+])
        (M0 g)[f-/+ of a line] = lm/lp[d] . g[line]
        (A5 f)[line]          += Dt[d] f[d, line]

    Raises ``NotFusable`` unless the four matrices are reproduced from
    these factors to ``tol``."""
    A1, M6, M0, A5 = (np.asarray(ops[k], dtype=float)
                      for k in ('A1', 'M6', 'M0', 'A5'))
    nf, nu = M0.shape
    n1 = int(round(nu**(1.0/nd)))

    if n1 < 2 or n1**nd != nu or nf != 2*nd*n1**(nd - 1):
        raise NotFusable('not a tensor-product point set')

    nl = nu // n1
    scale = lambda A: tol*max(1.0, np.abs(A).max())
    st = dict(n1=n1, nlines=nl, stride=[], base=[], fm=[], fp=[], Dg=[],
              Lg=[], lm=[], lp=[], Dt=[])
    used = np.zeros(nf, dtype=int)

    for d in range(nd):
        A1d, M6d = A1[d*nu:(d + 1)*nu], M6[d*nu:(d + 1)*nu]
        A5d = A5[:, d*nu:(d + 1)*nu]

        adj = (A1d != 0) | (A5d != 0)
        lines = _components(adj | adj.T)
        if len(lines) != nl or any(len(l) != n1 for l in lines):
            raise NotFusable('operators do not act along lines')

        strides = {l[i + 1] - l[i] for l in lines for i in range(n1 - 1)}
        if len(strides) != 1:
            raise NotFusable('lines are not arithmetic progressions')
        lines.sort()

        base, fms, fps = [], [], []
        ref = None
        for rows in lines:
            fs = np.flatnonzero(np.any(M6d[rows] != 0, axis=0))
            if len(fs) != 2:
                raise NotFusable('a line is not corrected from two points')

            # which end: the interpolated index coordinate of the point
            pos = [float(M0[f, rows] @ np.arange(n1)) for f in fs]
            fmn, fpl = (fs[0], fs[1]) if pos[0] < pos[1] else (fs[1], fs[0])
            used[[fmn, fpl]] += 1

            cur = (A1d[np.ix_(rows, rows)], M6d[np.ix_(rows, [fmn, fpl])],
                   M0[fmn, rows], M0[fpl, rows], A5d[np.ix_(rows, rows)])
            if ref is None:
                ref = cur
            elif any(np.abs(a - b).max() > scale(b)
                     for a, b in zip(cur, ref)):
                raise NotFusable('lines carry different 1-D operators')

            base.append(rows[0])
            fms.append(int(fmn))
            fps.append(int(fpl))

        st['stride'].append(int(strides.pop()))
        st['base'].append(base)
        st['fm'].append(fms)
        st['fp'].append(fps)
        for k, v in zip(('Dg', 'Lg', 'lm', 'lp', 'Dt'), ref):
            st[k].append(np.array(v))

    if np.any(used != 1):
        raise NotFusable('flux points are not the ends of the lines')

    # Rebuild the operators from the factors: everything outside the line
    # structure must vanish
    B1, B6, B0, B5 = (np.zeros_like(A) for A in (A1, M6, M0, A5))
    for d in range(nd):
        for b, fmn, fpl in zip(st['base'][d], st['fm'][d], st['fp'][d]):
            rows = b + st['stride'][d]*np.arange(n1)
            B1[np.ix_(d*nu + rows, rows)] = st['Dg'][d]
            B6[np.ix_(d*nu + rows, [fmn, fpl])] = st['Lg'][d]
            B0[fmn, rows], B0[fpl, rows] = st['lm'][d], st['lp'][d]
            B5[np.ix_(rows, d*nu + rows)] = st['Dt'][d]

    for A, B in ((A1, B1), (M6, B6), (M0, B0), (A5, B5)):
        if np.abs(A - B).max() > scale(A):
            raise NotFusable('operators have entries outside the lines')

    return st


def gradflux_tp_source(be, ops, tplargs, pts, LD, nthreads=None, rowcls=None,
                       affine=False, gather=False, dynamic=False):
    """Source of the sum-factorised fused kernel; arguments and return
    value as ``fused.gradflux_source``.

    ``gather``: the kernel forms the common solution at its flux points
    itself (``fusion.fold_conu``: the interior ``intconu`` launch and its
    pass over the flux-point array disappear).  Instead of one bulk copy of
    the block's ``ucomm`` tile every thread copies a few points, each from
    the address a per-block index table names -- the trace of whichever
    side of the interface a one-sided LDG flux takes (``sfp`` + index), or
    the block's own ``ucomm`` entry where another kernel (boundary,
    partition boundary) has stored the common value (index < 0) -- with
    per-thread asynchronous copies straight into shared memory.

    ``dynamic``: the blocks of the launch are handed out dynamically -- a
    CTA draws its next block from a device counter (``sched``) two
    iterations ahead, the last CTA to finish re-arms the counter.  Used
    where the element kernel is split into the blocks that touch a
    partition boundary and the rest, the latter running next to the halo
    exchange whose NCCL kernel holds some SMs for part of the time: CTAs
    that start late simply draw fewer blocks.

    Half blocks (``gradflux-split``, opt-in): where a whole element
    block fills an SM's shared memory (hexahedra, p = 4, fp64: 208 KB) the
    kernel works on half blocks instead -- the columns of ``C_SUB/2``
    elements, two co-resident CTAs of half the threads per SM -- so that
    the FP64-bound flux phase of one CTA overlaps the shared-memory-bound
    line phases of the other (measured with ``n-soa = 4``, where the
    *storage* has half-width blocks: -13 %, r02m; that setting slows every
    other kernel down).  The storage layout stays as it is: a half block's
    rows are 16-byte runs of the full rows, fetched with per-thread
    asynchronous copies instead of one bulk copy, and stored with the same
    16-byte stores as before."""
    nd, nv = tplargs['ndims'], tplargs['nvars']
    st = tp_structure(ops, nd)
    n1, nl = st['n1'], st['nlines']
    nu, nf = n1**nd, 2*nd*nl
    isz = np.dtype(be.fpdtype).itemsize
    gcsub, GLD = be.csubsz, LD                 # storage layout

    if LD != nv*gcsub:
        raise NotFusable('unexpected leading dimension')
    if gather and be.soasz != gcsub:
        raise NotFusable('gather form needs one SoA group per block')

    # Half blocks?
    linear = 'linear' in tplargs['ktype']
    geo_est = 4096
    fits = lambda ld: (227*1024) // ((nu + nf + nd*nu)*ld*isz + geo_est)
    SPLIT = 1
    if (getattr(be, 'gradflux_split', True) and linear and
        be.soasz == gcsub and gcsub % 2 == 0 and (gcsub // 2)*isz >= 16 and
        getattr(be, 'gradflux_groups', 1) == 1 and
        not getattr(be, 'gradflux_threads', 0) and nthreads is None and
        getattr(be, 'gradflux_maxctas', 2) >= 2 and
        fits(LD) == 1 and fits(LD // 2) >= 2):
        SPLIT = 2

    class _Half:
        """The backend as the kernel sees its (half-width) blocks."""
        def __init__(self, be, c):
            self._be, self.soasz, self.csubsz = be, c, c

        def __getattr__(self, k):
            return getattr(self._be, k)

    csub = gcsub // SPLIT
    LD = nv*csub
    gbe, be = be, (_Half(be, csub) if SPLIT > 1 else be)

    # Columns per work item: one 16-byte access
    NC = 16 // isz
    if LD % NC or be.soasz % NC:
        raise NotFusable('columns do not group into 16-byte accesses')
    NCG = LD // NC
    comps = 'xyzw'[:NC]
    vec = {(8, 2): 'double2', (4, 4): 'float4'}[isz, NC]

    affine = bool(affine and linear)

    # Occupancy plan: as many CTAs per SM as the shared-memory footprint
    # allows (two when a block is half an SM's worth, e.g. n-soa = 4 in
    # fp64), sharing a budget of 512 threads.  Co-resident CTAs work on
    # different blocks and drift apart, so the FP64-bound flux phase of one
    # overlaps the shared-memory-bound line phases of the other.
    smem_est = (nu + nf + nd*nu)*LD*isz + 4096
    nctas = max(1, min(getattr(be, 'gradflux_maxctas', 2),
                       (227*1024) // smem_est))
    if nthreads is None:
        nthreads = getattr(be, 'gradflux_threads', 0) or 512 // nctas

    # Warp groups: the elements of a block never interact inside this
    # kernel, so the CTA is split into NG groups of GT threads, each owning
    # the columns of C_SUB/NG elements and synchronising on its own named
    # barrier.  The groups run the same phases half a block out of step
    # (the second starts once the first has finished its first phase 1),
    # so the FP64-bound flux phase of one overlaps the shared-memory-bound
    # line phases of the other.
    NG = 1 if gather else getattr(be, 'gradflux_groups', 1)
    if (NG < 1 or csub % (NG*NC) or be.soasz != csub or nthreads % (32*NG)
            or (nthreads // NG) % (csub // NG)):
        NG = 1
    GT, H = nthreads // NG, csub // NG
    CPV = H // NC                              # column groups per variable
    NCGH = nv*CPV                              # ... and per warp group

    if nthreads % csub or GT < NCGH:
        raise NotFusable('thread count does not fit the block layout')

    NLG = GT // NCGH                           # line groups per round
    R = -(-nl // NLG)                          # rounds per direction
    ROWB = LD*isz
    nrounds = -(-nu*H // GT)

    defs = [('NDIMS', nd), ('NVARS', nv), ('NPTS', nu), ('NFPTS', nf),
            ('NVERTS', tplargs.get('nverts', 0)), ('NEED_RCPDJAC', 1),
            ('LD', LD), ('NTHREADS', nthreads), ('NROUNDS', nrounds),
            ('ROWB', ROWB), ('NCGH', NCGH), ('NLG', NLG), ('NLINES', nl),
            ('NG', NG), ('GT', GT), ('H', H), ('CPV', CPV),
            ('NGP', nf*csub), ('NGR', -(-nf*csub // nthreads)),
            ('SPLIT', SPLIT), ('GC_SUB', gcsub), ('GLD', GLD),
            ('GROWB', GLD*isz), ('CH16', csub*isz // 16)]
    defs += ph.physics_defines(tplargs['c'], tplargs.get('visc_corr', 'none'),
                               True)

    K = ConstPool(isz == 8)

    who = dict(acond='gtid < H', aelem='grp*H + gtid',
               mine='grp*H + gtid % H', lcond='gtid < NDIMS*H',
               lelem='grp*H + gtid % H', lcomp='gtid / H')
    geo = geometry_source(be, tplargs, pts, GT, affine, who=who)
    if affine and not geo['geo_post']:
        affine = False
    geom = geo['geom']

    # -- per-thread line descriptors (formed once per launch) ---------------
    # g_lines[d][line] = {first solution-point row, row of the '-' flux
    # point, row of the '+' flux point} as byte offsets; the low four bits
    # of the flux-point entries carry the row's need class
    if rowcls is not None and max(rowcls) < 16:
        cls = [int(c) for c in rowcls]
        fm_arg = ',\n         const int* __restrict__ fmask'
        fm_load = 'const unsigned fm = (unsigned) __ldg(fmask + rb);'
    else:
        cls, fm_arg, fm_load = None, '', ''

    # Packed: {first solution-point row | row of the '-' flux point << 16,
    # row of the '+' flux point | need classes of the two << 16, 24}.  A
    # thread keeps two words per work item for the whole launch; row
    # offsets and write masks are re-formed from them inside each phase
    # (a few integer operations) behind an optimisation barrier.  Without
    # it the compiler hoists the nine offsets and six class masks of a
    # thread out of the block loop and, at 128 registers, spills them:
    # their reloads sat on the critical path of every work item (ncu
    # r02c: 6.5 % of the kernel's stall samples on those ``LDL``).
    tab = []
    for d in range(nd):
        for b, fmn, fpl in zip(st['base'][d], st['fm'][d], st['fp'][d]):
            tab += [b | fmn << 16,
                    fpl | ((cls[fmn] | cls[fpl] << 8) << 16 if cls else 0)]
    tabsrc = (f'static __device__ __align__(8) const unsigned '
              f'g_lines[{len(tab)}] = {{{", ".join(map(str, tab))}}};')

    desc = []
    for d in range(nd):
        for r in range(R):
            desc.append(f'''
    unsigned dx{d}_{r} = 0, dy{d}_{r} = 0;
    const bool on{d}_{r} = lg + {r*NLG} < NLINES && lg < NLG;
    if (on{d}_{r})
    {{
        dx{d}_{r} = g_lines[{2*d*nl} + 2*(lg + {r*NLG})];
        dy{d}_{r} = g_lines[{2*d*nl} + 2*(lg + {r*NLG}) + 1];
    }}''')
    desc = ''.join(desc)

    def unpack(d, r, ends=False, masks=False, glob=False):
        """Row offsets (bytes, this item's columns included) of work item
        ``(d, r)``: ``lb`` first solution point, ``lm``/``lp`` its flux
        points; ``wm``/``wp`` whether those rows are needed; ``glob``: the
        same rows in the storage layout (``glb``, ``glm``, ``glp``)."""
        L = [f'    unsigned qx = dx{d}_{r}, qy = dy{d}_{r};',
             '    OPAQUE(qx); OPAQUE(qy);',
             '    const int lb = (int) (qx & 0xffffu)*ROWB + cb;']
        if ends:
            L += ['    const int lm = (int) (qx >> 16)*ROWB + cb, '
                  'lp = (int) (qy & 0xffffu)*ROWB + cb;']
        if glob:
            # (this item's columns in the storage layout; the half block's
            # own offset rides on the block pointers)
            L += ['    int cq = cb; OPAQUE(cq);',
                  f'    const int cbg = cq + (cq / (C_SUB*{isz}))'
                  f'*((GC_SUB - C_SUB)*{isz});']
        if glob and ends:
            L += ['    const int glm = (int) (qx >> 16)*GROWB + cbg, '
                  'glp = (int) (qy & 0xffffu)*GROWB + cbg;']
        elif glob:
            L += ['    const int glb = (int) (qx & 0xffffu)*GROWB + cbg;']
        if masks and cls:
            L += ['    const bool wm = (fm >> ((qy >> 16) & 0xffu)) & 1u, '
                  'wp = (fm >> (qy >> 24)) & 1u;']
        elif masks:
            L += ['    const bool wm = true, wp = true;']
        return L

    ld = lambda arr, off: (f'*reinterpret_cast<const fpvec_t *>({arr} + '
                           f'{off})')
    stv = lambda arr, off: f'*reinterpret_cast<fpvec_t *>({arr} + {off})'

    def lincomb(dst, terms, indent):
        """``dst.k = sum coef*src.k`` as FMA chains over the lanes."""
        out = []
        for k in comps:
            e = None
            for a, x in terms:
                if a == 0:
                    continue
                e = (f'{K(a)}*{x}.{k}' if e is None
                     else f'fma({K(a)}, {x}.{k}, {e})')
            out.append(f'{indent}{dst}.{k} = {e or "FP(0.0)"};')
        return out

    # Software pipelining of the line phases: the loads of work item k + 1
    # are written ahead of the arithmetic and the stores of item k.  The
    # compiler cannot make that move itself -- it has to assume that a
    # store into G may alias the next item's loads -- and without it a warp
    # alternates between waiting for shared memory and using the FP64
    # pipe (SASS of r02c: LDS x7, DFMA x70, STS x5 per direction).
    # (only where the loads of two work items fit the register file: at
    # p = 6 in fp32 -- nine float4 per item, three rounds -- the pipelined
    # form spills 112 bytes and runs 14.5 instead of 10.8 ms, r02s)
    swp = (getattr(be, 'gradflux_swp', True) and R == 1 and
           (n1 + 2)*NC*isz // 4 <= 28)

    def pipelined(items):
        """``items``: (condition, declarations, load lines, compute lines,
        lines every thread runs ahead of the compute step -- a barrier);
        the loads leave their values in the declared variables."""
        L = [l for it in items for l in it[1]]
        for k, (c, dcl, lds, cmp, pre) in enumerate(items):
            if k == 0 or not swp:
                L += [f'if ({c})', '{'] + lds + ['}']
            if swp and k + 1 < len(items):
                L += [f'if ({items[k + 1][0]})', '{'] + items[k + 1][2] + ['}']
            L += pre + [f'if ({c})', '{'] + cmp + ['}']
        return '\n        '.join(L)

    # -- phase 1: G[d] = Dg u + Lg (c-, c+) along the lines of d -------------
    p1 = []
    for d in range(nd):
        sb = st['stride'][d]*ROWB
        for r in range(R):
            t = f'{d}_{r}'
            dcl = ['fpvec_t ' + ', '.join(f'x{t}_{i}' for i in range(n1)) +
                   f', cm{t}, cp{t};']
            lds = unpack(d, r, ends=True)
            for i in range(n1):
                lds.append(f'    x{t}_{i} = {ld("Ub", f"lb + {i*sb}")};')
            lds += [f'    cm{t} = {ld("Cb", "lm")};',
                    f'    cp{t} = {ld("Cb", "lp")};']
            cmp = unpack(d, r) + ['    fpvec_t o;']
            for i in range(n1):
                terms = [(st['Dg'][d][i, j], f'x{t}_{j}') for j in range(n1)]
                terms += [(st['Lg'][d][i, 0], f'cm{t}'),
                          (st['Lg'][d][i, 1], f'cp{t}')]
                cmp += lincomb('o', terms, '    ')
                cmp.append(
                    f'    {stv("Gb", f"{d*nu*ROWB + i*sb} + lb")} = o;')
            p1.append((f'on{d}_{r}', dcl, lds, cmp, []))
    p1 = pipelined(p1)

    # -- phase 3: gradients at the flux points -> HBM -------------------------
    # Lines of direction a end on the two faces normal to a; all ndims
    # gradient components are interpolated along the line.  Affine
    # elements: G still holds the reference gradient and the (constant)
    # metric is applied to the interpolated values.
    metric_load = f'''
        // Constant metric of the {NC} elements this item's columns belong to
        // (QS was filled by the first threads of the block)
        // scaled by 1/|J|
        fpdtype_t sP[{NC}][NDIMS][NDIMS];
        UNROLL for (int k = 0; k < {NC}; k++)
        {{
            const fpdtype_t *q = QS + (e0 + k)*QSTRIDE + NDIMS*NDIMS + 1;
            UNROLL for (int i = 0; i < NDIMS; i++)
                UNROLL for (int j = 0; j < NDIMS; j++)
                    sP[k][i][j] = q[i*NDIMS + j];
        }}'''
    late = getattr(be, 'gradflux_metric_late', False)
    p3 = []
    for a in range(nd):
        sb = st['stride'][a]*ROWB
        for r in range(R):
            L = ([f'if (on{a}_{r})', '{'] +
                 unpack(a, r, ends=True, masks=True, glob=True))
            # One body per set of live ends: with a one-sided LDG flux a
            # line usually has one live end, and the interpolation to the
            # other one is not formed either
            for cond, ends in (('wm & wp', 'mp'), ('wm', 'm'), ('wp', 'p')):
                kw = 'if' if ends == 'mp' else 'else if'
                L += [f'    {kw} ({cond})', '    {']
                for d in range(nd):
                    L.append('        fpvec_t ' + ', '.join(
                        f't{e}{d}' for e in ends) + ';')
                    L.append('        {')
                    for i in range(n1):
                        L.append(
                            f'            const fpvec_t g{i} = '
                            f'{ld("Gb", f"{d*nu*ROWB + i*sb} + lb")};')
                    for e in ends:
                        L += lincomb(f't{e}{d}',
                                     [(st['l' + e][a][i], f'g{i}')
                                      for i in range(n1)], '            ')
                    L.append('        }')

                if affine and late:
                    # (the metric is fetched after the interpolation, so
                    # that its registers do not limit the loads in flight)
                    L.append(metric_load)
                for e in ends:
                    t = f't{e}'
                    L += ['        {', f'            char *vo = vfp + gl{e};']
                    for dp in range(nd):
                        if affine:
                            L.append('            { fpvec_t o;')
                            for ki, k in enumerate(comps):
                                ex = None
                                for d in range(nd):
                                    ex = (f'sP[{ki}][{d}][{dp}]*{t}{d}.{k}'
                                          if ex is None else
                                          f'fma(sP[{ki}][{d}][{dp}], '
                                          f'{t}{d}.{k}, {ex})')
                                L.append(f'            o.{k} = {ex};')
                            L.append(f'            {stv("vo", dp*nf*GLD*isz)}'
                                     ' = o; }')
                        else:
                            L.append(f'            {stv("vo", dp*nf*GLD*isz)} = '
                                     f'{t}{dp};')
                    L.append('        }')
                L.append('    }')
            L.append('}')
            p3.append('\n        '.join(L))
    p3 = '\n        '.join(p3)

    # -- phase 5: divergence --------------------------------------------------
    # in-place line transforms for all but the last direction ...
    def p5_loads(a, r):
        t = f'{a}_{r}'
        sb = st['stride'][a]*ROWB
        dcl = ['fpvec_t ' + ', '.join(f'f{t}_{i}' for i in range(n1)) +
               ';']
        lds = unpack(a, r)
        for i in range(n1):
            lds.append(f'    f{t}_{i} = '
                       f'{ld("Gb", f"{a*nu*ROWB + i*sb} + lb")};')
        return dcl, lds

    p5 = []
    for a in range(nd - 1):
        sb = st['stride'][a]*ROWB
        for r in range(R):
            t = f'{a}_{r}'
            dcl, lds = p5_loads(a, r)
            cmp = unpack(a, r) + ['    fpvec_t o;']
            for i in range(n1):
                cmp += lincomb('o', [(st['Dt'][a][i, j], f'f{t}_{j}')
                                     for j in range(n1)], '    ')
                cmp.append(
                    f'    {stv("Gb", f"{a*nu*ROWB + i*sb} + lb")} = o;')
            p5.append((f'on{a}_{r}', dcl, lds, cmp, []))

    # ... the last one in registers, summed with the others on the way out
    # (its own flux component is not touched by the transforms above and
    # is fetched ahead of the barrier that separates the two steps)
    a = nd - 1
    sb = st['stride'][a]*ROWB
    for r in range(R):
        t = f'{a}_{r}'
        dcl, lds = p5_loads(a, r)
        cmp = unpack(a, r, glob=True) + ['    fpvec_t o;']
        sbg = st['stride'][a]*GLD*isz
        for i in range(n1):
            cmp += lincomb('o', [(st['Dt'][a][i, j], f'f{t}_{j}')
                                 for j in range(n1)], '    ')
            for d in range(nd - 1):
                cmp.append(f'    {{ const fpvec_t y = '
                           f'{ld("Gb", f"{d*nu*ROWB + i*sb} + lb")};')
                cmp.append('      ' + ' '.join(f'o.{k} += y.{k};'
                                               for k in comps) + ' }')
            cmp.append(f'    {stv("fop", f"{i*sbg} + glb")} = o;')
        p5.append((f'on{a}_{r}', dcl, lds, cmp,
                   ['GSYNC();'] if (r == 0 and nd > 1) else []))
    p5 = pipelined(p5)

    # -- metric terms of the work items' columns (affine) ---------------------
    if affine:
        metric_p3 = '' if late else metric_load
        p2 = ''
        # (1/|J| rides on the viscosity: the viscous flux is linear in g)
        p4_xform = 'transform_grad(g, s, FP(1.0));'
        p4_visc = 'viscous_flux_add_sc(ureg[r], g, ft, rcpdjac_v);'
    else:
        metric_p3 = ''
        p2 = f'''
        // ---- phase 2: physical gradient (in place) ---------------------
        for (int item = gtid; item < NPTS*H; item += GT)
        {{
            const int e = grp*H + item % H, p = item / H;
            if (bq*C_SUB + e >= neles)
                continue;
{geom}
            fpdtype_t g[NDIMS][NVARS];
            UNROLL for (int d = 0; d < NDIMS; d++)
                UNROLL for (int v = 0; v < NVARS; v++)
                    g[d][v] = G[(d*NPTS + p)*LD + COFF(e, v, NVARS)];

            transform_grad(g, s, rcpdjac_v);

            UNROLL for (int d = 0; d < NDIMS; d++)
                UNROLL for (int v = 0; v < NVARS; v++)
                    G[(d*NPTS + p)*LD + COFF(e, v, NVARS)] = g[d][v];
        }}
        GSYNC();
'''
        p4_xform = '(void) rcpdjac_v;'
        p4_visc = 'viscous_flux_add(ureg[r], g, ft);'

    geo_words = geo['geo_words']
    smem = ((nu + nf + nd*nu)*LD + geo_words)*isz + 64
    if gather:
        smem += (nf*csub + nf + 4)*4
    smem += 16
    if smem > 227*1024:
        raise NotFusable(f'needs {smem} bytes of shared memory')

    stagger = getattr(be, 'gradflux_stagger', 1)
    if NG > 1:
        gsync = f'#define GSYNC() bar_sync_named(1 + grp, GT)'
        hold = '''
    // The second group starts once the first has finished a phase of its
    // first block (one thread polls, the rest wait on the group barrier)
    if (grp > 0 && (long long) blockIdx.x < nblk)
    {
        if (gtid == 0)
            flag_wait(flag);
        GSYNC();
    }
'''
        release = ('if (it == 0 && grp == 0 && gtid == 0) flag_set(flag);')
        issue = '''if (gtid == 0 && more)
        {
            // the group that arrives last starts the copy
            if (atomicAdd(cnt, 1) == NG - 1)
            {
                *reinterpret_cast<volatile int *>(cnt) = 0;
                fetch(nxt, it + 1);
            }
        }'''
    else:
        gsync = '#define GSYNC() __syncthreads()'
        hold, release = '', ''
        issue = '''if (tid == 0 && more)
            fetch(nxt, it + 1);'''
    rel1 = release if stagger == 1 else ''
    rel3 = release if stagger != 1 else ''

    from pyfr_b200.kernels.mul import _cpasync_src

    if dynamic:
        if SPLIT > 1 or NG > 1:
            raise NotFusable('dynamic blocks: whole blocks, one warp group')
        d_arg = ',\n         int* __restrict__ sched'
        # Blocks are drawn in chunks of DCH consecutive ones: chunk c of
        # the first gridDim.x belongs to CTA c, further chunks go to
        # whoever asks first.  NB[k & 1] holds the block of iteration k,
        # found during iteration k - 2; the ticket is drawn at the top of
        # that iteration and first touched after its first barrier (the
        # compiler parks it in local memory at once -- the kernel is at
        # its register limit -- which stalls warp 0 for the atomic's round
        # trip: 10 % of the kernel with one ticket per block, r02p2; hence
        # the chunks)
        DCH = 4
        defs.append(('DCH', DCH))
        dyn_first = '''
    int tk = 0;
    blk = (long long) blockIdx.x*DCH;
    if (tid == 0)
    {
        long long b1 = blk + 1;
        if (DCH == 1)
            b1 = (long long) (atomicAdd(sched, 1) + (int) gridDim.x)*DCH;
        NB[1] = (b1 < nblk) ? (int) b1 : -1;
    }
    __syncthreads();'''
        dyn_top = '''const int nxt = NB[(it + 1) & 1];
        const bool more = nxt >= 0;
        const bool draw = more && (nxt + 1) % DCH == 0;
        if (tid == 0 && draw)
            tk = atomicAdd(sched, 1);'''
        dyn_p1 = '''if (tid == 0)
        {
            long long b2 = draw ? (long long) (tk + (int) gridDim.x)*DCH
                                : nxt + 1;
            NB[it & 1] = (more && b2 < nblk) ? (int) b2 : -1;
        }'''
        dyn_p3 = ''
        dyn_next = 'blk = more ? nxt : nblk;'
        # ... the last CTA out re-arms the counters for the next launch
        dyn_last = '''
    if (tid == 0)
    {
        __threadfence();
        if (atomicAdd(sched + 1, 1) == (int) gridDim.x - 1)
        {
            sched[0] = 0; sched[1] = 0;
            __threadfence();
        }
    }'''
    else:
        d_arg = dyn_first = dyn_p1 = dyn_p3 = dyn_last = ''
        dyn_top = '''const long long nxt = bq + gridDim.x;
        const bool more = nxt < nblk;'''
        dyn_next = 'blk = nxt;'
    blkid = '#define BLOCK_ID(i) ((long long) (i))'

    split = SPLIT > 1
    nb_expr = 'nblocks*SPLIT' if split else 'nblocks'
    cpsrc = _cpasync_src if (gather or split) else ''

    # Whole rows by bulk copy: where the csub points of a flux-point row
    # take their values from one row of ``sfp`` (structured numbering:
    # every face but those whose neighbours sit one element further on)
    # the row is one contiguous run and a single TMA copy fetches it
    rows = (gather and not split and nf < nthreads and
            getattr(be, 'gather_rows', True))
    # per-thread copies complete on the block's mbarrier
    onbar = gather and not split

    if gather:
        g_arg = (',\n         const int* __restrict__ gidx,'
                 '\n         const fpdtype_t* __restrict__ sfp'
                 + (',\n         const int* __restrict__ growd'
                    if rows else ''))
        g_lambdas = f'''
    // Common solution by gather: this thread's points of a block (point =
    // flux-point row x element) and where each comes from ({"-2: fetched with its whole row; " if rows else ""}-1: the
    // block's own ucomm entry).  The indices of the next block are parked
    // in shared memory a phase ahead of the copies that need them (each
    // thread reads back what it copied itself)
    auto gidx_load = [&](long long b)
    {{
        const long long rbn = b / SPLIT;
        const int hf = (int) (b % SPLIT);
        UNROLL for (int r = 0; r < NGR; r++)
        {{
            const int p = tid + r*NTHREADS;
            if (p < NGP)
                cp_async<4>(GIX + p, gidx + rbn*(NFPTS*GC_SUB)
                            + (p / C_SUB)*GC_SUB + hf*C_SUB + p % C_SUB);
        }}
        {"""if (tid < NFPTS)
            cp_async<4>(ROWD + tid, growd + b*(NFPTS + 1) + tid);
        if (tid == 0)
            cp_async<4>(ROWD + NFPTS, growd + b*(NFPTS + 1) + NFPTS);""" if rows else ""}
    }};
    auto gather = [&](long long b, unsigned n)
    {{
        const long long rbn = b / SPLIT;
        const int hf = (int) (b % SPLIT);
        {"// (the indices were part of the previous block's barrier phase)" if onbar and not dynamic else ""}
        {"if (n == 0)" if onbar and not dynamic else ""}
        cp_async_wait_all();
        {"""// (the bulk copies of u, the vertices and ROWD[NFPTS] whole rows
        // complete on one mbarrier)
        if (tid == 0)
            fetch(b, n, ROWD[NFPTS]);
        if (tid < NFPTS && ROWD[tid] >= 0)
            tma_load_1d(C + tid*LD, sfp + ROWD[tid], ROWB, &bars[0]);""" if rows else ""}
        UNROLL for (int r = 0; r < NGR; r++)
        {{
            const int p = tid + r*NTHREADS;
            if (p < NGP && GIX[p] != -2)
            {{
                const int row = p / C_SUB, e = p % C_SUB, gi = GIX[p];
                const fpdtype_t *from = (gi >= 0) ? sfp + gi
                    : ucomm + rbn*ucomm_bsz + row*GLD + hf*C_SUB + e;
                fpdtype_t *to = C + row*LD + e;
                UNROLL for (int v = 0; v < NVARS; v++)
                    cp_async<{isz}>(to + v*C_SUB, from + v*GC_SUB);
            }}
        }}
        {"""// this thread's index slots are free again: the indices of the
        // block after b go in now, a whole iteration ahead of their use
        // (ncu r02q: fetched one phase ahead, 5 % of the kernel's stall
        // samples were this wait)
        if (b + gridDim.x < (long long) nblocks*SPLIT)
            gidx_load(b + gridDim.x);""" if not dynamic else ""}
        {"""// ... and once this thread's copies have landed they count as one
        // arrival on the block's barrier: nobody waits for them but the
        // threads that wait for the block (the cp.async.wait_all that stood
        // at the end of the loop held 4.4 % of the stall samples, r02s)
        cp_async_mbar_arrive(&bars[0]);""" if onbar else ""}
    }};'''
        # (dynamic blocks: the block after the next one is not known yet;
        # its indices are fetched at the top of the next iteration)
        g_top = 'if (more) gidx_load(nxt);' if dynamic else ''
        g_issue = 'if (more) gather(nxt, it + 1);'
    else:
        g_arg = g_lambdas = g_top = g_issue = ''

    if gather or split:
        g_first = ('''
    if (blk < nblk)
    {''' + ('''
        gidx_load(BLOCK_ID(blk));
        gather(BLOCK_ID(blk), 0);''' if gather else '') + ('''
        fetch(BLOCK_ID(blk), 0);''' if split else '') + '''
    }
    cp_async_wait_all();
    __syncthreads();''')
        g_wait = '' if onbar else 'cp_async_wait_all();'
    else:
        g_first = g_wait = ''

    if split:
        # Half blocks: 16-byte runs of the storage rows, one asynchronous
        # copy each, issued by all threads
        def strided(dst, srcp, bsz, nrows):
            return f'''
        {{
            const fpdtype_t *from = {srcp} + rbn*{bsz} + hf*C_SUB;
            for (int i = tid; i < ({nrows})*CH16; i += NTHREADS)
                cp_async16({dst} + (i / CH16)*C_SUB + (i % CH16)*{16 // isz},
                           from + (i / CH16)*GC_SUB + (i % CH16)*{16 // isz});
        }}'''

        fetch_body = ('const long long rbn = b / SPLIT;\n        '
                      'const int hf = (int) (b % SPLIT);'
                      + strided('U', 'u', 'u_bsz', 'NPTS*NVARS')
                      + ('' if gather else
                         strided('C', 'ucomm', 'ucomm_bsz', 'NFPTS*NVARS'))
                      + strided('(VSB + (n & 1)*V_WORDS)', 'verts',
                                'verts_bsz', 'NVERTS*NDIMS'))
        first_fetch = ''
        issue = '''if (more)
            fetch(nxt, it + 1);'''
        wait_tma = ''
    else:
        c_fetch = '' if gather else (
            '''tma_load_1d(C, ucomm + b*ucomm_bsz, C_WORDS*sizeof(fpdtype_t),
                    &bars[0]);''')
        fetch_body = f'''mbar_expect_tx(&bars[0], ({'U_WORDS' if gather else 'U_WORDS + C_WORDS'})*sizeof(fpdtype_t)
                                 {geo['geo_bytes']} + nrows*ROWB);
        tma_load_1d(U, u + b*u_bsz, U_WORDS*sizeof(fpdtype_t), &bars[0]);
        {c_fetch}
        {geo['geo_fetch']}'''
        first_fetch = '' if rows else '''if (tid == 0 && blk < nblk)
        fetch(BLOCK_ID(blk), 0);'''
        if rows:
            issue = ''
        wait_tma = 'mbar_wait(&bars[0], it & 1);'

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz, defs)}
typedef {vec} fpvec_t;
{_pipeline_src}{cpsrc}
{ph.flux_src}
{ph.visc_src}
{ph.geom_src}
{geo['gsrc']}
{tabsrc}
{K.decl()}

{blkid}
#define U_WORDS (NPTS*LD)
#define C_WORDS (NFPTS*LD)
#define G_WORDS (NDIMS*NPTS*LD)
#define V_WORDS (NVERTS*NDIMS*C_SUB)
{gsync}

// tensor-product element, {n1} points per line, {nl} lines per direction;
// {NG} warp group(s) of {GT} threads, each over {NCGH} column groups of {NC}
// columns x {NLG} line groups, {R} round(s) per direction{
    ', constant Jacobian' if affine else ''}
extern "C" __global__ void __launch_bounds__(NTHREADS, {nctas})
gradflux(int nblocks, int neles,
         const fpdtype_t* __restrict__ u, long long u_bsz,
         const fpdtype_t* ucomm, long long ucomm_bsz,
         fpdtype_t* vf, long long vf_bsz,
         fpdtype_t* __restrict__ fout, long long fout_bsz,
         {geo['gargs']}{fm_arg}{g_arg}{d_arg})
{{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    fpdtype_t *U = reinterpret_cast<fpdtype_t *>(smem_raw);
    fpdtype_t *C = U + U_WORDS;
    fpdtype_t *G = C + C_WORDS;
    {geo['geo_decl']}
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(
        G + G_WORDS + {geo_words});
    int *cnt = reinterpret_cast<int *>(bars + 1), *flag = cnt + 1;
    int *NB = cnt + 2, *GIX = NB + 2, *ROWD = GIX + NGP;
    (void) flag; (void) GIX; (void) ROWD; (void) NB;

    char *Ub = reinterpret_cast<char *>(U);
    char *Cb = reinterpret_cast<char *>(C);
    char *Gb = reinterpret_cast<char *>(G);

    const int tid = threadIdx.x;
    const int grp = tid / GT, gtid = tid % GT;

    // This thread's work items: column group cgl (of its warp group's
    // elements) of the lines lg, lg + NLG, ... of every direction
    const int cgl = gtid % NCGH, lg = gtid / NCGH;
    const int col0 = {'cgl*' + str(NC) if NG == 1 else
                      f'(cgl / CPV)*C_SUB + grp*H + (cgl % CPV)*{NC}'};
    const int cb = col0*{isz};
    const int e0 = (col0/(K_SOA*NVARS))*K_SOA + col0 % K_SOA;
    (void) e0;
{desc}

    // Stage the reference point set once per CTA
    {geo['geo_stage']}

    if (tid == 0)
    {{
        mbar_init(&bars[0], {'1 + NTHREADS' if onbar else '1'});
        cnt[0] = 0; cnt[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }}
    __syncthreads();

    auto fetch = [&](long long b, unsigned n, int nrows = 0)
    {{
        {fetch_body}
    }};
{g_lambdas}

    // (half) blocks of this launch
    const long long nblk = {nb_expr};
    long long blk = blockIdx.x;
{dyn_first}
    {first_fetch}
{g_first}
{hold}
    for (unsigned it = 0; blk < nblk; it++)
    {{
        // (block pointers are formed from the block number every time:
        // kept as induction variables they cost a dozen registers)
        long long bq = blk;
        OPAQUE64(bq);
        {dyn_top}
        // storage block and which half of its columns
        const long long rb = bq / SPLIT;
        const long long hoff = (bq % SPLIT)*C_SUB;
        char *vfp = reinterpret_cast<char *>(vf + rb*vf_bsz + hoff);
        char *fop = reinterpret_cast<char *>(fout + rb*fout_bsz + hoff);

        {wait_tma}
        {g_top}
        {geo['geo_blk']}
        {fm_load}
{geo['geo_elem']}

        // ---- phase 1: corrected transformed gradient ------------------
        {p1}

        // Keep the solution at this thread's flux-evaluation points
        fpdtype_t ureg[NROUNDS][NVARS];
        UNROLL for (int r = 0; r < NROUNDS; r++)
        {{
            const int item = gtid + r*GT;
            if (item < NPTS*H)
            {{
                const int e = grp*H + item % H, p = item / H;
                UNROLL for (int v = 0; v < NVARS; v++)
                    ureg[r][v] = U[p*LD + COFF(e, v, NVARS)];
            }}
        }}
        GSYNC();
        {rel1}

        // u and ucomm are consumed (by this group): once every group is
        // here the next block's copies are started behind the remaining
        // phases
        {issue}
        {g_issue}
        {dyn_p1}
{p2}
        // ---- phase 3: gradients at the flux points -> HBM ---------------
        {{
{metric_p3}
        {p3}
        }}
        GSYNC();
        {rel3}
        {dyn_p3}

        // ---- phase 4: transformed flux (in place over the gradient) -----
{geo['geo_post']}
        UNROLL for (int r = 0; r < NROUNDS; r++)
        {{
            const int item = gtid + r*GT;
            const int e = grp*H + item % H, p = item / H;
            if (item < NPTS*H && bq*C_SUB + e < neles)
            {{
{geom}
                fpdtype_t g[NDIMS][NVARS];
                UNROLL for (int d = 0; d < NDIMS; d++)
                    UNROLL for (int v = 0; v < NVARS; v++)
                        g[d][v] = G[(d*NPTS + p)*LD + COFF(e, v, NVARS)];

                {p4_xform}

                fpdtype_t ft[NDIMS][NVARS], fo[NDIMS][NVARS], pr, vel[NDIMS];
                inviscid_flux(ureg[r], ft, pr, vel);
                {p4_visc}
                transform_flux(ft, s, fo);

                UNROLL for (int d = 0; d < NDIMS; d++)
                    UNROLL for (int v = 0; v < NVARS; v++)
                        G[(d*NPTS + p)*LD + COFF(e, v, NVARS)] = fo[d][v];
            }}
        }}
        GSYNC();

        // ---- phase 5: divergence along the lines, summed -> HBM ----------
        {p5}
        {g_wait}
        GSYNC();
        {dyn_next}
    }}
{dyn_last}
}}
'''
    if nctas*(smem + 1024) > 227*1024:
        nctas = 1
    meta = dict(nthreads=nthreads, smem=smem, nctas=nctas, ngroups=NG,
                words_per_block=(2*nu + nf + nd*nf)*GLD, tensor=True,
                gather=gather, gather_points=nf*gcsub, split=SPLIT,
                gather_rows=bool(rows), dynamic=bool(dynamic),
                chunk=DCH if dynamic else 1)

    return src, 'gradflux', meta
