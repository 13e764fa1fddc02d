"""Graph-level kernel fusion.

The host code (the reference's system classes, or the mirror in
``pyfr_b200.host``) builds each RHS graph from individual kernels and, for
``blocks = True`` backends, marks the element-local chains that only
communicate through block-private temporaries with ``Graph.group(kerns,
subs)`` (``pyfr/solvers/baseadvecdiff/system.py:174-203,225-227``;
``pyfr/solvers/baseadvec/system.py:126-134``).  The reference's OpenMP
backend uses those hints to run a group block by block out of cache
(``pyfr/backends/openmp/types.py:136-186``); here they select generated
single-launch kernels that keep the temporaries in shared memory.

Every rewrite is checked structurally against the kernels it replaces and
silently declines when anything is unexpected -- the graph then simply runs
the individual (still GPU) kernels.
"""

import numpy as np

from pyfr_b200.kernels import fused as kfused
from pyfr_b200.kernels import fused_euler as keuler
from pyfr_b200.kernels import dense as kdense
from pyfr_b200.kernels import mul as kmul
from pyfr_b200.kernels import tensor as ktensor


def leaves(k):
    if hasattr(k, 'kernels'):
        return [l for c in k.kernels for l in leaves(c)]
    return [k]


def _same(a, b):
    """Do two matrix handles address the same storage window?"""
    return (a is b or (a.data == b.data and a.nrow == b.nrow and
                       a.ncol == b.ncol and a.blocksz == b.blocksz))


def _block_off(m):
    return getattr(m, 'ba', 0)


def _root(m):
    return getattr(m, 'parent', m)


def _rows_of(m):
    """(root matrix, first row, nrow) of a matrix or row/column slice."""
    return _root(m), getattr(m, 'ra', 0), m.nrow


def row_need_classes(be, mat, nrows, nblocks, maxclasses=15):
    """Which rows of ``mat`` (an element buffer whose first ``nrows`` rows
    are addressed by interface views; further row groups are reached
    through the views' row strides) are ever *read* through a view, per
    element block.

    Rows with the same need pattern over all blocks form a class.  Returns
    ``(cls, masks)``: ``cls[row]`` in ``[0, nclasses)`` and ``masks[block]``
    with bit ``c`` set when class ``c`` is read in that block -- or None
    when everything is read everywhere (or the pattern is too irregular
    to be worth encoding).  A one-sided LDG flux (|beta| = 1/2) reads the
    gradient on one side of each interface only, so on meshes whose blocks
    agree on which faces are left-hand sides about half of the rows are
    dead stores."""
    isz, LD = mat.itemsize, mat.leaddim
    lo = mat.offset // isz
    hi = lo + mat.nbytes // isz
    need = np.zeros((nblocks, nrows), dtype=bool)

    for v, mode in be.view_uses:
        if 'r' not in mode or not any(_root(m) is mat for m in v._mats):
            continue

        mp = v.mapping.get()[0].astype(np.int64)
        sel = (mp >= lo) & (mp < hi)
        idx = mp[sel] - lo
        blk, row = idx // mat.blocksz, (idx % mat.blocksz) // LD

        if v.nvrow > 1:
            rs = v.rstrides.get()[0][sel]
            if np.any(rs != nrows*LD):
                return None
        if np.any(row >= nrows) or np.any(blk >= nblocks):
            return None

        need[blk, row] = True

    if need.all():
        return None

    pats, cls = np.unique(need.T, axis=0, return_inverse=True)
    if len(pats) > maxclasses:
        return None

    masks = np.zeros(nblocks, dtype=np.int32)
    for c, pat in enumerate(pats):
        masks |= pat.astype(np.int32) << c

    return cls.ravel(), masks


def region_is_affine(verts, tol=1e-12):
    """Do all elements of a linear-element vertex matrix ``(nverts, ndims,
    neles)`` have a constant Jacobian?  True when every non-constant
    monomial coefficient of the multilinear map vanishes, i.e. the
    elements are parallelograms / parallelepipeds."""
    # ``verts`` may be the column slice of the linear region
    root = _root(verts)
    V = root.get()
    if root is not verts:
        nv = V.shape[1]
        V = V[..., verts.ca // nv:verts.cb // nv]
    nverts, nd, ne = V.shape

    if nverts != 2**nd or ne == 0:
        return False

    # Vertex n sits at (+-1, ..) with sign bit e of n along axis e; the
    # coefficient of monomial m is sum_n prod_{e in m} sign_e(n) V[n]/2^nd
    sg = np.array([[1.0 if (n >> e) & 1 else -1.0 for e in range(nd)]
                   for n in range(nverts)])
    scale = np.abs(V - V.mean(axis=0)).max()

    for m in range(1, 2**nd):
        axes = [e for e in range(nd) if (m >> e) & 1]
        if len(axes) < 2:
            continue
        w = np.prod(sg[:, axes], axis=1)
        if np.abs(np.einsum('n,nie->ie', w, V)).max() > tol*scale*nverts:
            return False

    return True



def conu_fold_plan(be, C):
    """Can the interior ``intconu`` that stores the common solution into
    ``C`` (the first ``nfpts`` rows of an element type's ``vect_fpts``,
    ``pyfr/solvers/baseadvecdiff/elements.py:42-46``) be folded into the
    element kernel that consumes it?

    With |ldg-beta| = 1/2 the common solution at an interior flux point is
    the trace of one of the two sides (``navstokes/kernels/intconu.mako``):
    nothing is computed, a value is moved.  The element kernel can fetch it
    from where ``disu`` left it -- the neighbour's (or its own) entry of
    ``scal_fpts`` -- while it stages its block, which saves the ``intconu``
    launch and its read and write of the flux-point array.  The rewrite
    reaches back into the graph committed just before the present one
    (``disu``, ``intconu``, packs; the reference's ``g1`` of
    ``baseadvecdiff/system.py:64-92``), which must not have run yet.

    Returns ``None`` or a dict with the graph, the kernel to drop, the
    trace matrix and the per-block gather indices (``-1``: the common value
    has been stored into ``C`` by a boundary or partition-boundary
    kernel)."""
    if not getattr(be, 'conu_fold', True):
        return None

    # (a flux-point row of one variable must be worth a copy: 16-byte rows
    # -- fp32 at n-soa = 4 -- were a loss, p = 6: 10.8 + 2.2 -> 16.3 ms, r02q)
    if be.csubsz*C.itemsize < 32:
        return None

    g = be.last_committed() if getattr(be, 'last_committed', None) else None
    if g is None or getattr(g, 'started', False):
        return None

    # On a partitioned mesh intconu is what the exchange of the traces
    # (issued by the same graph) overlaps with.  It can still go if the
    # element kernel takes its place: split into the blocks that touch a
    # partition boundary -- they need the halo, ``mpiconu`` stores their
    # common values in the present graph -- and the rest, which moves into
    # the earlier graph behind the exchange (``bnd`` below)
    overlap = any(w == 'xchg' for w, o in g.program)
    cur = getattr(be, '_fusing', None)
    if overlap and (not getattr(be, 'gradflux_overlap', True) or
                    cur is None):
        return None

    ks = [k for w, k in g.program
          if w == 'kernel' and getattr(k, 'kind', None) == 'intconu']
    if len(ks) != 1 or not ks[0].info.get('both'):
        return None

    k = ks[0]
    i = k.info
    beta = i['tplargs']['c']['ldg-beta']
    if abs(beta) != 0.5:
        return None

    VF = _root(C)
    nf, LD, isz = C.nrow, C.leaddim, C.itemsize
    csub, soa, nv = be.csubsz, be.soasz, i['tplargs']['nvars']
    views = [i[n] for n in ('ulin', 'urin', 'ulout', 'urout')]

    if any(v.nvrow != 1 or v.nvcol != nv or len(v._mats) != 1
           for v in views):
        return None

    S = views[0]._mats[0]
    if (_root(S) is not S or views[1]._mats[0] is not S or S.nrow != nf or
        S.leaddim != LD or S.nblocks != VF.nblocks or
        any(_root(v._mats[0]) is not VF or _rows_of(v._mats[0])[1] != 0
            for v in views[2:])):
        return None

    nblocks = VF.nblocks
    if S.blocksz*nblocks >= 2**31:
        return None

    src = (views[1] if beta > 0 else views[0]).mapping.get()[0]
    src = src.astype(np.int64) - S.offset // isz
    if len(src) and (src.min() < 0 or src.max() >= S.blocksz*nblocks):
        return None

    gidx = np.full((nblocks, nf*csub), -1, dtype=np.int32)
    for v in views[2:]:
        d = v.mapping.get()[0].astype(np.int64) - VF.offset // isz
        blk, rem = np.divmod(d, VF.blocksz)
        row, col = np.divmod(rem, LD)
        e = (col // (soa*nv))*soa + col % soa

        if len(d) and (d.min() < 0 or blk.max() >= nblocks or
                       row.max() >= nf or np.any(col % (soa*nv) >= soa) or
                       np.any(gidx[blk, row*csub + e] != -1)):
            return None

        gidx[blk, row*csub + e] = src

    # Rows whose csub points come from one row of the trace matrix, in
    # order: a single contiguous run (column 0 of a row, csub entries per
    # variable as in the destination).  rowd[block, row] = its offset, or
    # -1; rowd[block, nf] = how many such rows the block has
    g3 = gidx.reshape(nblocks, nf, csub)
    whole = ((g3[:, :, :1] >= 0) & (g3[:, :, :1] % LD == 0) &
             (g3 == g3[:, :, :1] + np.arange(csub))).all(axis=2)
    rowd = np.where(whole, g3[:, :, 0], -1).astype(np.int32)
    rowd = np.concatenate([rowd, whole.sum(axis=1, dtype=np.int32)[:, None]],
                          axis=1)

    bnd = None
    if overlap:
        mk = [o for w, o in cur.program
              if w == 'kernel' and getattr(o, 'kind', None) == 'mpiconu']
        if not mk:
            return None

        bnd = np.zeros(nblocks, dtype=bool)
        for o in mk:
            v = o.info['ulout']
            v = getattr(v, 'view', v)
            d = v.mapping.get()[0].astype(np.int64) - VF.offset // isz
            if len(v._mats) != 1 or _root(v._mats[0]) is not VF or \
               (len(d) and (d.min() < 0 or d.max() >= VF.blocksz*nblocks)):
                return None
            bnd[d // VF.blocksz] = True

        # (worth it only if most of the kernel can run behind the exchange;
        # the boundary launch takes the range of blocks up to the last one
        # on a partition boundary)
        if (np.flatnonzero(bnd).max() + 1) > 0.5*nblocks:
            return None

    return dict(graph=g, kernel=k, sfp=S, gidx=gidx, rowd=rowd, whole=whole,
                bnd=bnd)


def apply_conu_fold(plan, moved=()):
    """Drops the folded ``intconu`` from the graph that held it; ``moved``:
    kernels of the present graph that run at the end of that one instead
    (behind its exchange)."""
    g, k = plan['graph'], plan['kernel']
    g.program = [(w, o) for w, o in g.program if o is not k]
    g.program += [('kernel', m) for m in moved]
    g._plan()
    g.folded = getattr(g, 'folded', []) + [k]


def fuse_gradflux(be, kerns, subs):
    """tgradpcoru .. tdivtpcorf of one element type -> ``gradflux``."""
    from pyfr_b200.providers import B200Kernel

    if len(kerns) != 6:
        return None

    k0, k1, k2, k3, k4, k5 = kerns
    g2, g3, g4 = leaves(k2), leaves(k3), leaves(k4)

    if not (getattr(k0, 'kind', None) == 'mul' and
            getattr(k1, 'kind', None) == 'mul' and
            getattr(k5, 'kind', None) == 'mul' and
            all(k.kind == 'gradcoru' for k in g2) and
            all(k.kind == 'mul' for k in g3) and
            all(k.kind == 'tflux' for k in g4)):
        return None

    i0, i1, i5 = k0.info, k1.info, k5.info
    U, G, C, FOUT = i0['b'], i0['out'], i1['b'], i5['out']

    if (i0['beta'] != 0 or i0['alpha'] != 1 or i1['beta'] != 1 or
        i1['alpha'] != 1 or i5['beta'] != 0 or i5['alpha'] != 1 or
        not _same(i1['out'], G) or not _same(i5['b'], G)):
        return None

    tpl = g4[0].info['tplargs']
    nd, nv = tpl['ndims'], tpl['nvars']
    nu, nf = U.nrow, C.nrow
    LD = U.leaddim

    if (len(g3) != nd or G.nrow != nd*nu or len(g2) != len(g4) or
        any(k.info['tplargs'].get('shock_capturing', 'none') != 'none' or
            'fused' in k.info['tplargs']['ktype'] or not k.info['viscous']
            for k in g4)):
        return None

    # gradcoru_fpts: ndims multiplies by M0 from G[d] into vect_fpts[d]
    M0 = g3[0].info['A']
    VF = _root(g3[0].info['out'])
    for d, k in enumerate(g3):
        i = k.info
        if (i['beta'] != 0 or i['alpha'] != 1 or
            not np.array_equal(i['A'], M0) or
            _rows_of(i['b']) != (_root(G), d*nu, nu) or
            _rows_of(i['out']) != (VF, d*nf, nf)):
            return None

    if VF.nrow != nd*nf or VF.leaddim != LD or FOUT.leaddim != LD:
        return None

    isz = U.itemsize

    ops = dict(A1=i0['A'], M6=i1['A'], M0=M0, A5=i5['A'])
    out = []

    # The common solution gathered by the element kernel itself (all
    # regions of the element type or none)
    fold = conu_fold_plan(be, C) if be.gradflux_tensor else None
    fold_gidx = fold_rowd = None
    moved = []
    if fold is not None and fold['bnd'] is not None and len(g4) != 1:
        fold = None

    # Dead-store elimination on vect_fpts
    rneed = row_need_classes(be, VF, nf, VF.nblocks) if be.dead_rows else None

    # One launch per mesh region (curved / linear)
    for kg, kt in zip(g2, g4):
        ti, gi = kt.info, kg.info
        ktype = ti['tplargs']['ktype']

        if (gi['tplargs']['ktype'] != ktype or ti['dims'] != gi['dims'] or
            _root(ti['f']) is not _root(G) or _root(ti['u']) is not _root(U)
            or _root(gi['gradu']) is not _root(G)):
            return None

        npts, neles = ti['dims']
        b0 = _block_off(ti['u'])
        nblocks = -(-neles // be.csubsz)
        pts = ti['upts'].get() if ti['upts'] is not None else None

        affine = ('linear' in ktype and be.affine_fastpath and
                  region_is_affine(ti['verts']))
        # Tensor-product elements take the sum-factorised kernel; anything
        # else (or a structure the generator cannot verify) the
        # table-driven one
        src = None
        if be.gradflux_tensor:
            try:
                src, name, meta = ktensor.gradflux_tp_source(
                    be, ops, ti['tplargs'], pts, LD,
                    rowcls=None if rneed is None else rneed[0], affine=affine,
                    gather=fold is not None,
                    dynamic=fold is not None and fold['bnd'] is not None
                )
            except kfused.NotFusable:
                src = None
        if src is None:
            if fold is not None:
                # (the table-driven kernel has no gather form: start over
                # without the fold)
                be.conu_fold, keep = False, be.conu_fold
                try:
                    return fuse_gradflux(be, kerns, subs)
                finally:
                    be.conu_fold = keep
            src, name, meta = kfused.gradflux_source(
                be, ops, ti['tplargs'], pts, LD,
                rowcls=None if rneed is None else rneed[0], affine=affine
            )
        fn = be.pointwise._function(src, name)
        fn.set_smem(meta['smem'])

        off = lambda m: m.data + b0*m.blocksz*isz
        args = [('i', nblocks), ('i', neles),
                ('p', off(U)), ('l', U.blocksz),
                ('p', off(C)), ('l', C.blocksz),
                ('p', off(VF)), ('l', VF.blocksz),
                ('p', off(FOUT)), ('l', FOUT.blocksz)]

        if 'linear' in ktype:
            v = ti['verts']
            args += [('p', v.data), ('l', v.blocksz)]
            geo = [v]
        else:
            s, r = ti['smats'], gi['rcpdjac']
            args += [('p', s.data), ('l', s.blocksz), ('p', r.data),
                     ('l', r.blocksz)]
            geo = [s, r]

        words = meta['words_per_block']*nblocks
        if fold is not None:
            if fold_gidx is None:
                gi = fold['gidx']
                if meta['gather_rows']:
                    # points fetched with their whole row
                    gi = np.where(np.repeat(fold['whole'], be.csubsz, axis=1),
                                  -2, gi).astype(np.int32)
                fold_gidx = be.const_matrix(gi.reshape(1, -1),
                                            dtype=np.int32, tags={'noblock'})
                fold_rowd = be.const_matrix(fold['rowd'].reshape(1, -1),
                                            dtype=np.int32, tags={'noblock'})
            ngp = meta['gather_points']
            gargs = [('p', fold_gidx.data + b0*ngp*4),
                     ('p', fold['sfp'].data)]
            if meta['gather_rows']:
                gargs.append(('p', fold_rowd.data + b0*(nf + 1)*4))
            # (+ the index tables, one 32-bit word per point and per row)
            words += nblocks*(ngp + nf + 1)*4 // isz
        else:
            gargs = []
        if rneed is not None:
            fm = be.const_matrix(rneed[1][None, b0:b0 + nblocks],
                                 dtype=np.int32, tags={'noblock'})
            args.append(('p', fm.data))
            geo = geo + [fm]

            # Rows actually written
            cls, masks = rneed
            live = sum(int(((masks[b0:b0 + nblocks] >> c) & 1).sum()) *
                       int((cls == c).sum()) for c in range(cls.max() + 1))
            words -= nd*(nf*nblocks - live)*LD

        args += gargs
        if fold is not None:
            geo = geo + [fold_gidx, fold_rowd, fold['sfp']]

        info = dict(replaces=kerns, dead_rows=rneed is not None,
                    affine=affine, tensor=bool(meta.get('tensor')),
                    gather=fold is not None, split=meta.get('split', 1),
                    gidx=fold_gidx, rows=meta.get('gather_rows'),
                    rowd=fold_rowd if fold is not None else None)

        if fold is not None and fold['bnd'] is not None:
            # Two launches over block ranges, blocks drawn dynamically:
            # [0, nbb) holds every block on a partition boundary (those
            # elements come first in a partition, ``pyfr/partitioners/
            # base.py:286-290``) and runs here; the rest runs at the end of
            # the earlier graph, behind its exchange
            nbb = int(np.flatnonzero(fold['bnd'][b0:b0 + nblocks]).max()) + 1
            for part, lo, hi in (('interior', nbb, nblocks),
                                 ('boundary', 0, nbb)):
                if hi <= lo:
                    continue
                sch = be.const_matrix(np.zeros((1, 2), dtype=np.int32),
                                      dtype=np.int32, tags={'noblock'})

                # block-indexed arguments start at block lo
                pargs, step = [('i', hi - lo),
                               ('i', neles - lo*be.csubsz)], None
                for (c, v), nxt in zip(args[2:], args[3:] + [(None, 0)]):
                    if c == 'p' and nxt[0] == 'l':
                        v += lo*nxt[1]*isz
                    pargs.append((c, v))
                # (fmask, gidx, growd: one entry / NGP / nf + 1 per block)
                tail = []
                if rneed is not None:
                    tail.append(('p', fm.data + lo*4))
                tail.append(('p', gargs[0][1] + lo*ngp*4))
                tail.append(gargs[1])
                if meta['gather_rows']:
                    tail.append(('p', gargs[2][1] + lo*(nf + 1)*4))
                pargs = pargs[:len(pargs) - len(tail)] + tail
                pargs.append(('p', sch.data))

                kern = B200Kernel(
                    be, fn, (min(-(-(hi - lo) // meta['chunk']),
                                 be.sm_count*meta['nctas']), 1, 1),
                    (meta['nthreads'], 1, 1), meta['smem'], pargs,
                    mats=[U, C, VF, FOUT, G, sch] + geo, misc=[meta],
                    traffic=words*isz*(hi - lo)//nblocks, kind='gradflux',
                    info=dict(info, part=part, nblocks=hi - lo)
                )
                (moved if part == 'interior' else out).append(kern)
            continue

        # (a kernel working on half blocks walks twice as many)
        ngrid = min(nblocks*meta.get('split', 1), be.sm_count*meta['nctas'])
        out.append(B200Kernel(
            be, fn, (ngrid, 1, 1),
            (meta['nthreads'], 1, 1), meta['smem'], args,
            mats=[U, C, VF, FOUT, G] + geo, misc=[meta],
            traffic=words*isz, kind='gradflux', info=info
        ))

    if fold is not None:
        if fold['bnd'] is not None and not out:
            # (no block touches a partition boundary: keep one launch here)
            out, moved = moved, []
        apply_conu_fold(fold, moved)

    return out


def _rk_tail(be, kerns, fout):
    """Splits an optional trailing ``rkvdh2`` kernel off a group: returns
    (remaining kernels, rk template arguments or None, extra kernel
    arguments, extra matrices, rk traffic in bank passes).  The stage
    update must consume the RHS bank ``fout`` as its ``r2``."""
    if not kerns or getattr(kerns[-1], 'kind', None) != 'rkvdh2':
        return kerns, None, [], [], 0, None
    if not be.rk_fusion:
        raise kfused.NotFusable('rk fusion disabled')

    i = kerns[-1].info
    tpl, r1 = i['tplargs'], i['r1']
    if not _same(i['r2'], fout) or r1.traits != fout.traits:
        raise kfused.NotFusable('rkvdh2 does not follow this RHS')

    from pyfr_b200.providers import rt_scalar

    regs = [r1] + ([i['rold'], i['rerr']] if tpl['errest'] else [])
    args = [a for m in regs for a in (('p', m.data), ('l', m.blocksz))]
    dtarg, dtset = rt_scalar(be)
    args.append(dtarg)

    last = tpl['stage'] == tpl['nstages'] - 1
    passes = 2 + (2 if tpl['errest'] else 0) - (1 if last else 0)
    return kerns[:-1], tpl, args, regs, passes, dtset


def _bind_dt(k, dtset):
    """Gives a fused kernel the ``bind(dt=)`` of the rkvdh2 it absorbed."""
    k.rtnames, k.bind = ('dt',), lambda dt=0.0: dtset(dt)


def fuse_tdivtconf_negdivconf(be, kerns, subs):
    """tdivtconf (out += M3 @ scal_fpts) followed by negdivconf."""
    from pyfr_b200.providers import B200Kernel

    if len(kerns) not in (2, 3) or getattr(kerns[0], 'kind', None) != 'mul':
        return None

    allk = kerns
    kerns, rk, rkargs, rkmats, rkpasses, dtset = _rk_tail(
        be, kerns, kerns[0].info['out'])
    if len(kerns) != 2:
        return None

    km, kn = kerns
    ln = leaves(kn)
    if not (getattr(km, 'kind', None) == 'mul' and len(ln) == 1 and
            ln[0].kind == 'negdivconf'):
        return None

    im, ineg = km.info, ln[0].info
    out, b = im['out'], im['b']

    if (not _same(ineg['tdivtconf'], out) or ineg['tplargs']['src_macros'] or
        ineg['dims'][0] != out.nrow):
        return None

    LD, nv = b.leaddim, ineg['tplargs']['nvars']
    nblocks = -(-b.ncol // LD)
    r = ineg['rcpdjac']

    isz = b.itemsize
    if be.dense_mul and not rk and kdense.is_dense(im['A'], LD, isz):
        gen = (kdense.dense_mma_source if be.dense_mma and isz == 8
               else kdense.dense_mul_source)
        src, name, meta = gen(be, im['A'], LD, im['alpha'], im['beta'],
                              negdiv_nvars=nv)
        ngrid = min(-(-nblocks // meta['nb']), be.sm_count*meta['nctas'])
    else:
        src, name, meta = kmul.mul_source(
            be, im['A'], LD, im['alpha'], im['beta'],
            # (eight row groups where the blocks are a full warp wide; on
            # narrow blocks they cost the second CTA per SM: p = 6 fp32 at
            # n-soa = 4, 4.4 -> 9.5 ms, r02q)
            smem_budget=be.smem_budget,
            rowgroups=be.mul_rowgroups or (8 if LD >= 32 else 4),
            negdiv_nvars=nv, rk=rk
        )
        ngrid = min(nblocks, be.sm_count*meta['nctas'])
    fn = be.pointwise._function(src, name)
    fn.set_smem(meta['smem'])

    isz = b.itemsize
    traffic = ((b.nrow + out.nrow*(2 if im['beta'] else 1) +
                out.nrow*rkpasses)*LD + out.nrow*be.csubsz)*nblocks*isz

    args = [('i', nblocks), ('p', b.data), ('l', b.blocksz), ('p', out.data),
            ('l', out.blocksz), ('p', r.data), ('l', r.blocksz)] + rkargs
    k = B200Kernel(
        be, fn, (ngrid, 1, 1), (meta['nthreads'], 1, 1),
        meta['smem'], args, mats=[b, out, r] + rkmats, misc=[meta],
        traffic=traffic,
        kind='mul+negdivconf+rkvdh2' if rk else 'mul+negdivconf',
        info=dict(replaces=allk)
    )

    # negdivconf carries the (unused here) run-time argument t
    k.rtnames = ()
    if rk:
        _bind_dt(k, dtset)
    return [k]


def fuse_fluxdiv(be, kerns, subs):
    """tdisf, tdivtpcorf, tdivtconf, negdivconf of one element type of an
    advection (Euler) system -> ``fluxdiv``."""
    from pyfr_b200.providers import B200Kernel

    if len(kerns) not in (4, 5) or not be.euler_fusion or \
       getattr(kerns[1], 'kind', None) != 'mul':
        return None

    allk = kerns
    kerns, rk, rkargs, rkmats, rkpasses, dtset = _rk_tail(
        be, kerns, kerns[1].info['out'])
    if len(kerns) != 4:
        return None

    k0, k1, k2, k3 = kerns
    g0, g3 = leaves(k0), leaves(k3)

    if not (all(k.kind == 'tflux' and not k.info['viscous'] for k in g0) and
            getattr(k1, 'kind', None) == 'mul' and
            getattr(k2, 'kind', None) == 'mul' and
            len(g3) == 1 and g3[0].kind == 'negdivconf'):
        return None

    i1, i2, i3 = k1.info, k2.info, g3[0].info
    F, FOUT, C = i1['b'], i1['out'], i2['b']

    if (i1['beta'] != 0 or i1['alpha'] != 1 or i2['beta'] != 1 or
        i2['alpha'] != 1 or not _same(i2['out'], FOUT) or
        not _same(i3['tdivtconf'], FOUT) or i3['tplargs']['src_macros']):
        return None

    tpl = g0[0].info['tplargs']
    nd = tpl['ndims']
    nu, nf = FOUT.nrow, C.nrow
    U = _root(g0[0].info['u'])
    LD = U.leaddim

    if (F.nrow != nd*nu or U.nrow != nu or C.leaddim != LD or
        FOUT.leaddim != LD or F.leaddim != LD or
        any(_root(k.info['f']) is not _root(F) or
            _root(k.info['u']) is not U for k in g0)):
        return None

    isz = U.itemsize
    ops = dict(A5=i1['A'], M3=i2['A'])
    out = []

    for kt in g0:
        ti = kt.info
        ktype = ti['tplargs']['ktype']
        npts, neles = ti['dims']
        b0 = _block_off(ti['u'])
        nblocks = -(-neles // be.csubsz)
        pts = ti['upts'].get() if ti['upts'] is not None else None

        src, name, meta = keuler.fluxdiv_source(be, ops, ti['tplargs'], pts,
                                                LD, rk=rk)
        fn = be.pointwise._function(src, name)
        fn.set_smem(meta['smem'])

        off = lambda m: m.data + b0*m.blocksz*isz
        args = [('i', nblocks), ('i', neles),
                ('p', off(U)), ('l', U.blocksz),
                ('p', off(C)), ('l', C.blocksz),
                ('p', off(FOUT)), ('l', FOUT.blocksz)]

        if 'linear' in ktype:
            v = ti['verts']
            args += [('p', v.data), ('l', v.blocksz)]
            geo = [v]
        else:
            s, r = ti['smats'], i3['rcpdjac']
            args += [('p', s.data), ('l', s.blocksz),
                     ('p', r.data + b0*r.blocksz*isz), ('l', r.blocksz)]
            geo = [s, r]

        if rk:
            # the stage registers of this region's blocks
            # (the trailing pointer is the run-time scalar dt)
            args += [(c, v + b0*rkmats[i // 2].blocksz*isz)
                     if c == 'p' and i < 2*len(rkmats) else (c, v)
                     for i, (c, v) in enumerate(rkargs)]

        kern = B200Kernel(
            be, fn, (min(nblocks, be.sm_count*meta['nctas']), 1, 1),
            (meta['nthreads'], 1, 1), meta['smem'], args,
            mats=[U, C, FOUT, F] + geo + rkmats, misc=[meta],
            traffic=(meta['words_per_block'] + rkpasses*nu*LD)*nblocks*isz,
            kind='fluxdiv+rkvdh2' if rk else 'fluxdiv',
            info=dict(replaces=allk)
        )
        if rk:
            _bind_dt(kern, dtset)
        out.append(kern)

    return out


_group_fusers = [fuse_gradflux, fuse_fluxdiv, fuse_tdivtconf_negdivconf]


def fuse_group(be, kerns, subs):
    for f in _group_fusers:
        try:
            new = f(be, kerns, subs)
        except (KeyError, AttributeError, AssertionError,
                kfused.NotFusable):
            new = None

        if new:
            return new

    return None


def elide_copy_fpts(be, program):
    """Drop ``copy_fpts`` when the common-solution kernels that follow it
    can store every flux point themselves.

    With |ldg-beta| = 1/2 the reference seeds the common solution with a
    full copy of the interpolated solution and lets ``intconu`` overwrite
    one side only (``pyfr/solvers/baseadvecdiff/elements.py:42-46``,
    ``navstokes/kernels/intconu.mako``).  Every flux point belongs to exactly
    one interface side, so when the con_u kernels of the graph cover the
    whole array the copy is a wasted pass; ``intconu`` is regenerated to
    store both sides."""
    kerns = [k for w, k in program if w == 'kernel']
    copies = [k for k in kerns if getattr(k, 'kind', None) == 'copy']
    conus = [k for k in kerns if getattr(k, 'kind', None) in
             ('intconu', 'mpiconu')]

    if not copies or not conus:
        return program

    # One copy per element type (mixed meshes: several); the interface
    # views span all of them
    pairs = [(cp.info['dst'], cp.info['src']) for cp in copies]
    base_d = {int(d.basedata) for d, s in pairs}
    base_s = {int(s.basedata) for d, s in pairs}
    if len(base_d) != 1 or len(base_s) != 1:
        return program

    if any(kerns.index(k) < kerns.index(cp) for k in conus for cp in copies):
        return program

    def span(m):
        lo = m.offset // m.itemsize
        return lo, lo + (m.nblocks - 1)*m.blocksz + m.nrow*m.leaddim

    dspan = [span(d) for d, s in pairs]
    sspan = [span(s) for d, s in pairs]

    def locate(mp, spans, mats):
        """(matrix number, block, offset in block) of view points"""
        mp = mp.astype(np.int64)
        which = np.full(len(mp), -1)
        blk, off = np.zeros_like(mp), np.zeros_like(mp)
        for j, ((lo, hi), m) in enumerate(zip(spans, mats)):
            sel = (mp >= lo) & (mp < hi)
            which[sel] = j
            blk[sel] = (mp[sel] - lo) // m.blocksz
            off[sel] = (mp[sel] - lo) % m.blocksz
        return which, blk, off

    # Coverage: interior kernels touch 2n points, partition-boundary ones n
    npts = 0
    for k in conus:
        i = k.info
        sides = [(i['ulin'], i['ulout'])]
        if k.kind == 'intconu':
            sides.append((i['urin'], i['urout']))

        for vin, vout in sides:
            vin, vout = getattr(vin, 'view', vin), getattr(vout, 'view', vout)

            # The trace and the common solution must be addressed alike
            if (int(vout.basedata) not in base_d or
                int(vin.basedata) not in base_s):
                return program

            wi, bi, oi = locate(vin.mapping.get()[0], sspan,
                                [s for d, s in pairs])
            wo, bo, oo = locate(vout.mapping.get()[0], dspan,
                                [d for d, s in pairs])
            if not (np.all(wi >= 0) and np.array_equal(wi, wo) and
                    np.array_equal(bi, bo) and np.array_equal(oi, oo)):
                return program

            npts += vin.n

    # Partition-boundary points are stored by ``mpiconu`` kernels, which
    # write their output for every beta (navstokes/kernels/mpiconu.mako)
    # but run in a later graph, after the halo has arrived: count the
    # points of every other view the backend has bound for *writing* into
    # the common-solution buffers
    here = {id(getattr(v, 'view', v)) for k in conus
            for v in (k.info['ulout'], k.info['urout']) if v is not None}
    seen = set()
    for v, mode in be.view_uses:
        if mode != 'w' or id(v) in here or id(v) in seen or \
           int(v.basedata) not in base_d:
            continue
        seen.add(id(v))

        if v.n:
            w, _, _ = locate(v.mapping.get()[0], dspan, [d for d, s in pairs])
            # (views into other buffers of the same allocation -- the
            # common fluxes in scal_fpts -- are not ours to count)
            if np.all(w < 0):
                continue
            if np.any(w < 0):
                return program

        npts += v.n

    total = 0
    for d, s in pairs:
        nele = s.ioshape[-1] if hasattr(s, 'ioshape') else None
        if nele is None:
            return program
        total += s.nrow*nele
    if npts != total:
        return program

    # Regenerate intconu with both-side stores
    repl = {}
    for k in conus:
        if k.kind == 'intconu' and not k.info['both']:
            i = k.info
            repl[k] = be.pointwise._conu(False, i['tplargs'], i['dims'],
                                         i['ulin'], i['urin'], i['ulout'],
                                         i['urout'], both=True)

    out = []
    for w, k in program:
        if any(k is cp for cp in copies):
            continue
        out.append((w, repl.get(k, k)))

    return out


# -- one launch for the per-neighbour kernels of a graph -------------------------
_sig_re = None


def _batched_source(src):
    """Turns a 1-D pointwise kernel into a device function and wraps it in
    a kernel that takes its arguments from a table indexed by
    ``blockIdx.y``.  Returns (source, kernel name, [parameter types])."""
    import re

    global _sig_re
    if _sig_re is None:
        _sig_re = re.compile(
            r'extern "C" __global__ void\s*(__launch_bounds__\([^)]*\))?\s*'
            r'(\w+)\s*\(([^)]*)\)', re.S
        )

    m = _sig_re.search(src)
    bounds, name, params = m.group(1) or '', m.group(2), m.group(3)
    params = [p.strip() for p in params.split(',') if p.strip()]
    types = [re.sub(r'\w+$', '', p).replace('__restrict__', '').strip()
             for p in params]

    impl = (f'static __device__ __forceinline__ void {name}_impl('
            + ', '.join(params) + ')')
    body = src[:m.start()] + impl + src[m.end():]

    fields = '\n'.join(f'    {t} a{i};' for i, t in enumerate(types))
    call = ', '.join(f'a.a{i}' for i in range(len(types)))
    wrap = f'''
// The same kernel over several argument sets (one per neighbouring
// partition): grid.y selects the set, grid.x covers the largest
struct batch_args_t
{{
{fields}
}};

extern "C" __global__ void {bounds}
{name}(const batch_args_t* __restrict__ tbl)
{{
    const batch_args_t a = tbl[blockIdx.y];
    {name}_impl({call});
}}
'''
    return body + wrap, name, types


def _arg_table(kerns):
    """The argument sets of ``kerns`` laid out as the C compiler lays out
    ``batch_args_t`` (natural alignment), as an int64 matrix."""
    import ctypes as ct

    size = {'p': 8, 'l': 8, 'd': 8, 'i': 4, 'f': 4}
    codes = kerns[0].argcodes
    offs, off = [], 0
    for c in codes:
        off = -(-off // size[c])*size[c]
        offs.append(off)
        off += size[c]
    stride = -(-off // 8)*8

    buf = np.zeros((len(kerns), stride), dtype=np.uint8)
    for j, k in enumerate(kerns):
        for c, o, v in zip(codes, offs, k._vals):
            raw = ct.string_at(ct.addressof(v), size[c])
            buf[j, o:o + size[c]] = np.frombuffer(raw, dtype=np.uint8)

    return buf.view(np.int64)


_batchable = ('pack', 'mpiconu', 'mpicflux')


def batch_launches(be, program):
    """Kernels a graph holds once per neighbouring partition -- ``pack``,
    ``mpiconu``, ``mpicflux`` (``pyfr/solvers/base/system.py:185-202``,
    ``pyfr/solvers/navstokes/inters.py:58-67``) -- are launches of one
    device function over independent argument sets: runs of them become a
    single launch whose ``blockIdx.y`` picks the argument set from a
    device-resident table.  At eight ranks this removes eight of the
    twelve per-neighbour launches of an RHS (10-25 us each)."""
    from pyfr_b200.providers import B200Kernel

    def kind_of(k):
        if not isinstance(k, B200Kernel) or k.smem or \
           getattr(k, 'rtnames', None):
            return None
        kind = k.kind or k.fn.name
        return kind if kind in _batchable else None

    out, run = [], []

    def flush():
        """``run``: consecutive kernels of one kind (with the exchange
        requests issued between them).  They act on different neighbours'
        points, so they may be regrouped by device function (the kernels
        of a one-sided LDG flux come in two variants, by rank parity)."""
        ks = [k for w, k in run if w == 'kernel']
        groups = {}
        for k in ks:
            groups.setdefault((id(k.fn), k.block), []).append(k)

        for g in groups.values():
            if len(g) == 1:
                out.append(('kernel', g[0]))
                continue

            src, name, types = _batched_source(g[0].fn.src)
            fn = be.pointwise._function(src, name)
            tbl = be.const_matrix(_arg_table(g), dtype=np.int64,
                                  tags={'noblock'})
            out.append(('kernel', B200Kernel(
                be, fn, (max(k.grid[0] for k in g), len(g), 1),
                g[0].block, 0, [('p', tbl.data)],
                mats=[m for k in g for m in k.mats] + [tbl],
                views=[v for k in g for v in k.views],
                traffic=sum(k.traffic for k in g),
                kind=g[0].kind or g[0].fn.name, info=dict(batched=g)
            )))

        out.extend(e for e in run if e[0] != 'kernel')
        run.clear()

    for w, obj in program:
        if w == 'kernel':
            kk = kind_of(obj)
            cur = next((kind_of(k) for ww, k in run if ww == 'kernel'), None)
            if kk is None or (cur is not None and kk != cur):
                flush()
            if kk is None:
                out.append((w, obj))
            else:
                run.append((w, obj))
        else:
            # exchanges between the members of a run stay behind them
            (run if run else out).append((w, obj))

    flush()
    return out
