"""Generator for the fused element kernel of the Navier-Stokes RHS.

One launch replaces the chain the reference runs as six to eight kernels
(``pyfr/solvers/baseadvecdiff/system.py:94-205``; the block-fusion group it
hands to ``Graph.group`` at ``:174-203``):

    tgradpcoru_upts   G  = (M4 - M6*M0) @ u
    tgradcoru_upts    G += M6 @ ucomm
    gradcoru_upts     G  = J^-T G                      (physical gradient)
    gradcoru_fpts     vect_fpts[d] = M0 @ G[d]
    tdisf             G  = S (Fi(u) + Fv(u, G))        (transformed flux)
    tdivtpcorf        fout = (M1 - M3*M2) @ G

``G`` (``ndims*nupts`` rows) never leaves the SM: per element block the
kernel reads ``u`` and the common solution and writes only the gradients at
the flux points and the partial divergence -- ``(2 nupts + nfpts +
ndims nfpts)*LD`` words instead of the ``~35 nupts*LD`` the unfused chain
moves.

How the operators are applied (sm_100a, FP64/FP32 FMA pipes):

* The operators of tensor-product elements are unions of small dense
  blocks: the rows belonging to one line of solution points read the same
  few inputs with the *same* coefficients as every other line.  The
  generator discovers this from the matrices alone: rows with identical
  column support form a *group*; groups whose coefficient blocks agree (up
  to a row/column permutation) form a *class*.  Each class becomes one
  short loop whose coefficients are immediates and whose row indices come
  from a small table, so the instruction footprint is a few KB instead of
  the hundreds of KB of a fully unrolled product, every lane is busy
  (lanes run over (group, column) pairs) and each input is read from
  shared memory once per group instead of once per row.
* ``(M1 - M3*M2)`` is block diagonal over lines in every direction, so the
  divergence is formed by in-place line transforms of the flux followed by
  one pass that sums the directions and streams the result to HBM.
* ``u`` and ``ucomm`` arrive by TMA bulk copy.  Both buffers are dead once
  phase 1 has run (each thread keeps the few ``u`` values its flux points
  need in registers), so the copies for the *next* block are issued right
  then and overlap phases 2-5.
* pointwise phases: a thread owns (point, element) pairs -- exactly the
  arithmetic of the stand-alone ``gradcoru``/``tflux`` kernels.
"""

import itertools as it

import numpy as np

from pyfr_b200.kernels import physics as ph
from pyfr_b200.kernels.mul import _pipeline_src


class NotFusable(Exception):
    pass


def _support(A, m):
    return tuple(np.flatnonzero(A[m]))


def _row_groups(terms, rows, smax=16):
    """Partition ``rows`` into groups whose inputs are contained in one
    common small input set (per term).  Rows are absorbed into the group
    with the smallest covering support, so a row with a structurally zero
    coefficient still joins its line."""
    sups = {m: tuple(frozenset(_support(A, m)) for A in terms) for m in rows}
    order = sorted(rows, key=lambda m: -sum(len(s) for s in sups[m]))
    groups = []            # [support per term, rows]

    for m in order:
        if sum(len(s) for s in sups[m]) > smax:
            groups.append([sups[m], [m]])
            continue

        fits = [g for g in groups
                if all(a <= b for a, b in zip(sups[m], g[0]))]
        if fits:
            min(fits, key=lambda g: sum(len(s) for s in g[0]))[1].append(m)
        else:
            groups.append([sups[m], [m]])

    return [sorted(g[1]) for g in groups]


def _canonical(blocks):
    """Canonical form of a group's coefficient blocks under a common row
    permutation and independent column permutations.  Returns
    (key, row order, [column order per block])."""
    n = blocks[0].shape[0]
    perms = it.permutations(range(n)) if n <= 6 else [tuple(range(n))]
    best = None

    for perm in perms:
        cols, parts = [], []
        for B in blocks:
            Bp = B[list(perm)]
            order = sorted(range(B.shape[1]), key=lambda j: tuple(Bp[:, j]))
            cols.append(order)
            parts.append(Bp[:, order])

        key = tuple(tuple(P.ravel()) for P in parts)
        if best is None or key < best[0]:
            best = (key, perm, cols)

    return best


class OpClass:
    """Groups sharing one coefficient pattern."""

    def __init__(self, nout, nins, coefs):
        self.nout, self.nins, self.coefs = nout, nins, coefs
        self.members = []       # (out rows, [input rows per term])


def build_classes(terms, max_classes=24, rows=None):
    """Decompose ``out = sum_t A_t @ src_t`` (restricted to the output
    ``rows`` when given) into line classes."""
    M = terms[0].shape[0]
    classes = {}

    for g in _row_groups(terms, range(M) if rows is None else rows):
        sups = [sorted(set().union(*[_support(A, m) for m in g]))
                for A in terms]
        blocks = [A[np.ix_(g, s)] for A, s in zip(terms, sups)]
        key, perm, cols = _canonical(blocks)

        rows = [g[i] for i in perm]
        ins = [[s[j] for j in order] for s, order in zip(sups, cols)]

        sig = (len(g), tuple(len(s) for s in sups), key)
        if sig not in classes:
            coefs = [B[list(perm)][:, order]
                     for B, order in zip(blocks, cols)]
            classes[sig] = OpClass(len(g), [len(s) for s in sups], coefs)

        classes[sig].members.append((rows, ins))

    if len(classes) > max_classes:
        raise NotFusable(f'{len(classes)} operator classes')

    return list(classes.values())


class ConstPool:
    """Operator coefficients.  A 64-bit immediate cannot be encoded in a
    DFMA, so literals cost two uniform-register moves per use; placed in
    ``__constant__`` memory they become constant-bank operands of the FMA
    itself.  fp32 coefficients stay literals (FFMA takes a 32-bit
    immediate)."""

    def __init__(self, use_table):
        self.use_table, self.vals = use_table, {}

    def __call__(self, a):
        a = float(a)
        if not self.use_table:
            return ph.fpconst(a)
        return f'KC[{self.vals.setdefault(a, len(self.vals))}]'

    def decl(self):
        if not self.vals:
            return ''
        body = ', '.join(ph.fpconst(v) for v in self.vals)
        return f'__constant__ fpdtype_t KC[{len(self.vals)}] = {{{body}}};'


def _fma_chain(pairs, K, acc=None):
    expr = acc
    for a, x in pairs:
        if a == 0:
            continue
        if expr is None:
            expr = f'{K(a)}*{x}'
        else:
            expr = f'fma({K(a)}, {x}, {expr})'
    return expr or 'FP(0.0)'


class PhaseEmitter:
    """Renders class loops and collects their index tables.

    Table entries are byte offsets (``row*LD*sizeof``) held as 32-bit
    integers, four to a 16-byte record, staged in shared memory once per
    CTA: a work item fetches its indices with a few 128-bit loads and
    forms every address with a single add."""

    def __init__(self, LD, itemsize, K, ncol=1):
        self.LD, self.isz, self.K, self.ncol = LD, itemsize, K, ncol
        self.tables = []        # (name, flat list of ints)
        self.staged = None      # names of the tables copied to smem

    def plan(self, budget):
        """Choose the tables staged in shared memory (smallest first)
        within ``budget`` bytes; the rest are read through L1."""
        self.staged, used = set(), 0
        for name, vals in sorted(self.tables, key=lambda t: len(t[1])):
            if used + 4*len(vals) <= budget:
                self.staged.add(name)
                used += 4*len(vals)

    @property
    def table_bytes(self):
        return sum(4*len(t) for n, t in self.tables if n in self.staged)

    def decls(self):
        out = []
        for name, vals in self.tables:
            flat = ', '.join(map(str, vals))
            out.append(f'static __device__ __align__(16) const int '
                       f'g_{name}[{len(vals)}] = {{{flat}}};')
        return '\n'.join(out)

    def smem_layout(self, base):
        """Carve the staged tables out of the smem pointer ``base``."""
        out, off = [], 0
        for name, vals in self.tables:
            if name in self.staged:
                out.append(f'const int *{name} = reinterpret_cast<int *>('
                           f'{base}) + {off};')
                off += len(vals)
            else:
                out.append(f'const int *{name} = g_{name};')
        return '\n    '.join(out)

    def stage(self, base):
        out, off = [], 0
        for n, v in self.tables:
            if n in self.staged:
                out.append(f'for (int i = tid; i < {len(v)}; i += NTHREADS) '
                           f'(reinterpret_cast<int *>({base}) + {off})[i] = '
                           f'g_{n}[i];')
                off += len(v)
        return '\n    '.join(out)

    def emit_planes(self, tag, blocks, boffs, extra, store):
        """``out[p] = sum_d blocks[d][p, :] @ src[boffs[d] + :] + extra``
        with one thread per (plane, column): the plane's outputs are
        accumulated in registers while each input is read from shared
        memory exactly once -- per output ``ndirs`` loads instead of the
        ``2*ndirs`` loads + stores of line-by-line in-place transforms.

        ``boffs``: first row of each direction's block in ``G``;
        ``extra(byte offset expr, at)``: optional additional addend read
        at the output point; ``store(byte offset expr, value)``."""
        LD, isz, K = self.LD, self.isz, self.K
        planes = find_planes(blocks)
        if planes is None:
            raise NotFusable('planes too large')

        # Classes of planes with identical coefficient blocks
        classes = {}
        for P in planes:
            key = tuple(tuple(B[np.ix_(P, P)].ravel()) for B in blocks)
            classes.setdefault(key, []).append(P)

        out = []
        for ci, (key, members) in enumerate(classes.items()):
            np_ = len(members[0])
            npad = -(-np_ // 4)*4
            coefs = [np.array(k).reshape(np_, np_) for k in key]

            tab = []
            for P in members:
                tab += [r*LD*isz for r in P] + [0]*(npad - np_)
            name = f'tab_{tag}_{ci}'
            self.tables.append((name, tab))

            at = lambda arr, off: (f'*reinterpret_cast<fpdtype_t *>('
                                   f'reinterpret_cast<char *>({arr}) + {off})')
            idx = lambda j: f'q.{"xyzw"[j % 4]}'
            ldq = lambda q: (f'const int4 q = *reinterpret_cast<const int4 *>'
                             f'({name} + g*{npad} + {4*q});')

            L = [f'for (int item = tid; item < {len(members)}*LD; '
                 'item += NTHREADS)', '{',
                 '    const int g = item / LD;',
                 f'    const int cb = (item - g*LD)*{isz};',
                 '    fpdtype_t ' + ', '.join(f'a{i}' for i in range(np_))
                 + ';']

            # Indices are fetched four at a time next to their use (and
            # again for the stores) rather than held in registers
            acc = [None]*np_
            for j in range(np_):
                if j % 4 == 0:
                    L.append('    {')
                    L.append('    ' + ldq(j // 4))
                for d, (C, bo) in enumerate(zip(coefs, boffs)):
                    if not np.any(C[:, j]):
                        continue
                    L.append(f'    const fpdtype_t x{d}_{j} = '
                             f'{at("G", f"{idx(j)} + cb + {bo*LD*isz}")};')
                    for i in range(np_):
                        if C[i, j] != 0:
                            a = K(C[i, j])
                            L.append(
                                f'    a{i} = {a}*x{d}_{j};'
                                if acc[i] is None else
                                f'    a{i} = fma({a}, x{d}_{j}, a{i});'
                            )
                            acc[i] = True
                if j % 4 == 3 or j == np_ - 1:
                    L.append('    }')

            for i in range(np_):
                if i % 4 == 0:
                    L.append('    {')
                    L.append('    ' + ldq(i // 4))
                off = f'{idx(i)} + cb'
                val = f'a{i}' if acc[i] else 'FP(0.0)'
                if extra is not None:
                    val = f'{val} + {extra(off, at)}'
                L.append('    ' + store(off, val, at))
                if i % 4 == 3 or i == np_ - 1:
                    L.append('    }')

            L.append('}')
            out.append('\n        '.join(L))

        return '\n        '.join(out)

    def emit(self, tag, classes, srcs, store, inplace=False, outtag=None,
             rep=None, vec=False):
        """``srcs[t]``: name of the (shared) array term ``t`` reads;
        ``store(byte offset expr, val)`` renders a store.  ``rep = (n,
        out stride, in stride)`` applies the same classes ``n`` times with
        all output / input rows shifted by the given byte strides (a block
        diagonal operator with identical blocks needs one table).

        ``outtag(row)``: class of an output row; the kernel holds a
        per-block mask ``fm`` and skips outputs whose class bit is clear.
        Lines are sub-divided by the classes of their outputs, so the test
        is a block-uniform branch on a compile-time bit, not per-lane
        arithmetic.

        A work item handles one line for ``self.ncol`` columns ``LD/ncol``
        apart: the index fetches and address additions are shared, the
        extra columns being reached through immediate offsets.

        ``vec``: a work item handles two *adjacent* columns with one
        16-byte access per row (rows are ``LD*isz`` bytes apart and ``LD`` is
        even, so every access stays aligned): half the loads, stores,
        index fetches and address additions per value for the same number
        of shared-memory wavefronts -- for phases that are bound by
        instruction issue rather than by shared-memory bandwidth."""
        LD, isz, K, out = self.LD, self.isz, self.K, []
        vec = bool(vec) and LD % 2 == 0
        NC = 2 if vec else (self.ncol if LD % self.ncol == 0 else 1)
        W, WB = LD // NC, (LD // NC)*isz
        CW = 2*isz if vec else isz          # bytes between items' columns

        if inplace:
            # A work item overwrites its own inputs after reading them
            # all; that is only an in-place *transform* if no other item
            # reads those rows (block-diagonal operator, one group per
            # block).  Dense blocks that were split into several groups
            # (more than smax inputs) do not qualify.
            owner = {r: (ci, mi) for ci, c in enumerate(classes)
                     for mi, (rows, ins) in enumerate(c.members)
                     for r in rows}
            for ci, c in enumerate(classes):
                for mi, (rows, ins) in enumerate(c.members):
                    if any(owner.get(k, (ci, mi)) != (ci, mi)
                           for k in ins[0]):
                        raise NotFusable('in-place transform over '
                                         'overlapping groups')

        for ci, c in enumerate(classes):
            nidx = (0 if inplace else c.nout) + sum(c.nins)
            npad = -(-nidx // 4)*4

            # Sub-divide by the output classes (one group when untagged)
            parts = {}
            for rows, ins in c.members:
                if inplace and not set(rows) <= set(ins[0]):
                    raise NotFusable('in-place transform needs rows within '
                                     'inputs')
                key = tuple(outtag(r) for r in rows) if outtag else ()
                parts.setdefault(key, []).append((rows, ins))

            for pi, (key, members) in enumerate(parts.items()):
                tab = []
                for rows, ins in members:
                    ent = [] if inplace else [r*LD*isz for r in rows]
                    for sidx in ins:
                        ent += [k*LD*isz for k in sidx]
                    tab += ent + [0]*(npad - nidx)

                name = f'tab_{tag}_{ci}_{pi}'
                self.tables.append((name, tab))
                nmem = len(members)

                L = []
                if key:
                    bits = ' | '.join(f'{1 << k}u' for k in sorted(set(key)))
                    L += [f'if (fm & ({bits}))']
                if rep:
                    L += [f'for (int item = tid; item < {rep[0]*nmem*W}; '
                          'item += NTHREADS)', '{',
                          f'    const int rd = item / {nmem*W};',
                          f'    const int ritem = item - rd*{nmem*W};',
                          f'    const int g = ritem / {W};',
                          f'    const int cbi = (ritem - g*{W})*{CW} + '
                          f'rd*{rep[2]};',
                          f'    const int cb = (ritem - g*{W})*{CW} + '
                          f'rd*{rep[1]};']
                else:
                    L += [f'for (int item = tid; item < {nmem*W}; '
                          'item += NTHREADS)', '{',
                          f'    const int g = item / {W};',
                          f'    const int cb = (item - g*{W})*{CW};',
                          '    const int cbi = cb;']

                for q in range(npad // 4):
                    L.append(f'    const int4 q{q} = *reinterpret_cast<const '
                             f'int4 *>({name} + g*{npad} + {4*q});')

                idx = lambda j: f'q{j // 4}.{"xyzw"[j % 4]}'
                at = lambda arr, off: (
                    f'*reinterpret_cast<fpdtype_t *>('
                    f'reinterpret_cast<char *>({arr}) + {off})')
                at2 = lambda arr, off: (
                    f'*reinterpret_cast<fpdtype2_t *>('
                    f'reinterpret_cast<char *>({arr}) + {off})')

                base = 0 if inplace else c.nout
                regs, off = [], base
                for t, n in enumerate(c.nins):
                    for j in range(n):
                        L.append(f'    const int i{t}_{j} = {idx(off + j)} '
                                 '+ cbi;')
                        if vec:
                            L.append(f'    const fpdtype2_t x{t}_{j} = '
                                     f'{at2(srcs[t], f"i{t}_{j}")};')
                            L.append(f'    const fpdtype_t x{t}_{j}_0 = '
                                     f'x{t}_{j}.x, x{t}_{j}_1 = x{t}_{j}.y;')
                            continue
                        for k in range(NC):
                            L.append(
                                f'    const fpdtype_t x{t}_{j}_{k} = '
                                f'{at(srcs[t], f"i{t}_{j} + {k*WB}")};')
                    regs.append([f'x{t}_{j}' for j in range(n)])
                    off += n

                for i in range(c.nout):
                    if inplace:
                        # Output i lives where the same point's input was
                        rows0, ins0 = members[0]
                        j = ins0[0].index(rows0[i])
                        if any(ins[0].index(rows[i]) != j
                               for rows, ins in members):
                            raise NotFusable('inconsistent in-place slots')
                        dst = f'i0_{j}'
                    else:
                        dst = f'{idx(i)} + cb'

                    body, vals = [], []
                    for k in range(NC):
                        pairs = [(c.coefs[t][i, j], f'{regs[t][j]}_{k}')
                                 for t in range(len(c.nins))
                                 for j in range(c.nins[t])]
                        vals.append(_fma_chain(pairs, K))
                        if not vec:
                            body.append(store(f'{dst} + {k*WB}', vals[-1],
                                              at))
                    if vec:
                        body.append(store(
                            dst, f'fpdtype2_t{{{vals[0]}, {vals[1]}}}', at2))

                    if key:
                        L.append(f'    if (fm & {1 << key[i]}u)')
                        L.append('    {')
                        L += ['        ' + b for b in body]
                        L.append('    }')
                    else:
                        L += ['    ' + b for b in body]

                L.append('}')
                out.append('\n        '.join(L))

        return '\n        '.join(out)


def find_planes(blocks, maxsize=36):
    """Partition the points into the connected components of the coupling
    graph of the square operator ``blocks`` (one per direction): for a
    tensor-product element and two directions these are its coordinate
    planes.  Returns a list of sorted index lists, or None if a component
    exceeds ``maxsize`` points (too many accumulators for one thread)."""
    n = blocks[0].shape[0]
    adj = sum((B != 0).astype(int) for B in blocks)
    adj = (adj + adj.T) != 0
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

        if len(comp) > maxsize:
            return None
        comps.append(sorted(comp))

    return comps


def geometry_source(be, tplargs, pts, nthreads, affine=False, who=None):
    """Source fragments that give the element kernels their metric terms
    (shared by ``gradflux_source`` and ``tensor.gradflux_tp_source``).

    Three forms: affine linear elements (one Jacobian per element, formed
    once per block and kept in registers), other linear elements (Jacobian
    from per-element monomial coefficients) and curved elements (``smats``
    and ``rcpdjac`` read from memory).  The fragments refer to the kernel's
    ``tid``, ``blk``, ``it``, ``bars``, ``G``/``G_WORDS`` and, inside a
    point loop, ``e`` (element of the block) and ``p`` (solution point);
    ``geom`` leaves ``s[NDIMS][NDIMS]`` and ``rcpdjac_v`` in scope.
    Returns a dict of strings plus ``geo_words`` (shared-memory words).

    ``who``: which threads form the per-element quantities and which
    element a thread works on, as C expressions (default: the first
    ``C_SUB`` / ``NDIMS*C_SUB`` threads of the CTA, element ``tid %
    C_SUB``); a kernel whose warp groups own disjoint element sets passes
    its own."""
    w = dict(acond='tid < C_SUB', aelem='tid', mine='tid % C_SUB',
             lcond='tid < NDIMS*C_SUB', lelem='tid % C_SUB',
             lcomp='tid / C_SUB')
    w.update(who or {})
    nd = tplargs['ndims']
    nu = len(pts) if pts is not None else 0
    isz = np.dtype(be.fpdtype).itemsize
    csub = be.csubsz
    linear = 'linear' in tplargs['ktype']

    geo_elem = geo_post = ''
    if linear:
        nverts = tplargs['nverts']
        gargs = 'const fpdtype_t* __restrict__ verts, long long verts_bsz'
        mj = (ph.multilinear_jacobian(tplargs['jac_exprs'], nd, nverts)
              if getattr(be, 'gradflux_monojac', True) else None)

        if mj is not None and affine and nthreads % csub == 0:
            # Constant Jacobian: only the constant monomial survives
            monos, W = mj
            k0 = monos.index(())
            jl = []
            for d in range(nd):
                for i in range(nd):
                    terms = ' '.join(
                        f'{"+" if W[d, k0, n] > 0 else "-"} '
                        f'{ph.fpconst(abs(W[d, k0, n]))}*'
                        f'VS[{n}*(NDIMS*C_SUB) + COFF(ea, {i}, NDIMS)]'
                        for n in range(nverts) if W[d, k0, n] != 0)
                    jl.append(f'jm[{d}][{i}] = {terms or "FP(0.0)"};')

            gsrc = (f'static __device__ const fpdtype_t c_pts[2] = '
                    '{FP(0.0), FP(0.0)};\n' + ph.smats_from_jac_src(nd))
            geo_elem = (r'''
        // Affine elements: one Jacobian per element.  C_SUB threads form
        // the metric terms of the block's elements; after the next barrier
        // every thread picks up those of the one element it works on
        // (tid % C_SUB, the same in every round) and keeps them in
        // registers through phases 2 and 4.
        if (''' + w['acond'] + r''')
        {
            const int ea = ''' + w['aelem'] + r''';
            fpdtype_t jm[NDIMS][NDIMS], sm[NDIMS][NDIMS], djac;
            ''' + '\n            '.join(jl) + r'''
            smats_detj_from_jac(jm, sm, djac);
            // (second copy: the metric scaled by 1/|J|, as the
            // flux-point gradients use it)
            const fpdtype_t rj = FP(1.0)/djac;
            UNROLL for (int i = 0; i < NDIMS; i++)
                UNROLL for (int j = 0; j < NDIMS; j++)
                {
                    QS[ea*QSTRIDE + i*NDIMS + j] = sm[i][j];
                    QS[ea*QSTRIDE + NDIMS*NDIMS + 1 + i*NDIMS + j] = rj*sm[i][j];
                }
            QS[ea*QSTRIDE + NDIMS*NDIMS] = rj;
        }
''')
            geo_post = r'''
        fpdtype_t sA[NDIMS][NDIMS], rjA;
        {
            const fpdtype_t *q = QS + (''' + w['mine'] + r''')*QSTRIDE;
            UNROLL for (int i = 0; i < NDIMS; i++)
                UNROLL for (int j = 0; j < NDIMS; j++)
                    sA[i][j] = q[i*NDIMS + j];
            rjA = q[NDIMS*NDIMS];
        }
'''
            geom = r'''
            const fpdtype_t (&s)[NDIMS][NDIMS] = sA;
            const fpdtype_t rcpdjac_v = rjA;
'''
            # Per-element stride of QS: the elements of a block are read
            # side by side with 16-byte loads, so consecutive elements
            # must start four banks apart modulo a permutation -- in fp64
            # a stride of 2 (mod 4) doubles (20 put elements e and e + 4
            # on the same banks: 29 M excess wavefronts, 8 % of the
            # kernel's shared-memory traffic, ncu r02s)
            qst = 2*(nd*nd + 1)
            while isz == 8 and qst % 4 != 2:
                qst += 2
            npt_words, q_words = 2, qst*csub
            gsrc = f'#define QSTRIDE {qst}\n' + gsrc
        elif mj is not None:
            # Per element: Q[q][i] = sum_n W[d][k][n] V[n][i] for every
            # (d, monomial k) with a non-zero coefficient; per point: the
            # monomial values.  j[d][i] = sum_k mono_k(x_p) Q[(d,k)][i].
            monos, W = mj
            used = [k for k in range(len(monos)) if monos[k] and
                    np.any(W[:, k])]
            nm = len(used) + len(used) % 2          # padded to pairs

            # Q is stored element by element, (entry, component) pairs
            # contiguous, so a lane fetches two values per 16-byte load;
            # the per-element stride is padded to keep the eight lanes of
            # a row group on distinct banks
            qidx, qsrc = {}, []
            for d in range(nd):
                for k in range(len(monos)):
                    if np.any(W[d, k]):
                        qidx[d, k] = len(qidx)
            nqv = len(qidx)*nd
            nqv += nqv % 2
            qstride = nqv + 2

            for (d, k), q in qidx.items():
                terms = ' '.join(
                    f'{"+" if W[d, k, n] > 0 else "-"} '
                    f'{ph.fpconst(abs(W[d, k, n]))}*v[{n}]'
                    for n in range(nverts) if W[d, k, n] != 0)
                qsrc.append(f'QS[e*{qstride} + {q}*NDIMS + i] = {terms};')

            vals = []
            for row in pts:
                mv = [np.prod(row[list(monos[k])]) for k in used]
                vals += mv + [0.0]*(nm - len(mv))
            rows = ', '.join(ph.fpconst(v) for v in vals)
            gsrc = (f'static __device__ const fpdtype_t c_pts[{len(pts)*nm}]'
                    f' = {{{rows}}};\n' + ph.smats_from_jac_src(nd))

            jl = []
            for d in range(nd):
                for i in range(nd):
                    e = None
                    for k in range(len(monos)):
                        if (d, k) not in qidx:
                            continue
                        q = f'qv[{qidx[d, k]*nd + i}]'
                        if not monos[k]:
                            e = q if e is None else f'({e} + {q})'
                        else:
                            m = f'mv[{used.index(k)}]'
                            e = (f'{m}*{q}' if e is None
                                 else f'fma({m}, {q}, {e})')
                    jl.append(f'jm[{d}][{i}] = {e or "FP(0.0)"};')

            geom = (rf'''
            fpdtype_t s[NDIMS][NDIMS], djac;
            {{
                fpdtype_t jm[NDIMS][NDIMS], qv[{nqv}], mv[{nm}];
                UNROLL for (int a = 0; a < {nqv // 2}; a++)
                {{
                    const fpdtype2_t t = reinterpret_cast<const fpdtype2_t *>(
                        QS + e*{qstride})[a];
                    qv[2*a] = t.x; qv[2*a + 1] = t.y;
                }}
                UNROLL for (int a = 0; a < {nm // 2}; a++)
                {{
                    const fpdtype2_t t = reinterpret_cast<const fpdtype2_t *>(
                        PTS + p*{nm})[a];
                    mv[2*a] = t.x; mv[2*a + 1] = t.y;
                }}
                ''' + '\n                '.join(jl) + r'''
                smats_detj_from_jac(jm, s, djac);
            }
            const fpdtype_t rcpdjac_v = FP(1.0)/djac;
''')
            npt_words = nu*nm
            q_words = qstride*csub
            geo_elem = (r'''
        // Element-wise Jacobian coefficients from the vertices
        if (''' + w['lcond'] + r''')
        {
            const int e = ''' + w['lelem'] + r''', i = ''' + w['lcomp'] + r''';
            fpdtype_t v[NVERTS];
            UNROLL for (int n = 0; n < NVERTS; n++)
                v[n] = VS[n*(NDIMS*C_SUB) + COFF(e, i, NDIMS)];
            ''' + '\n            '.join(qsrc) + '''
        }
''')
        else:
            gsrc = ph.linear_smats_src(nd, nverts, tplargs['jac_exprs'])
            rows = ', '.join(ph.fpconst(v) for row in pts for v in row)
            gsrc = (f'static __device__ const fpdtype_t c_pts[{len(pts)*nd}]'
                    f' = {{{rows}}};\n' + gsrc)
            geom = r'''
            fpdtype_t V[NVERTS][NDIMS], x[NDIMS], s[NDIMS][NDIMS], djac;
            UNROLL for (int n = 0; n < NVERTS; n++)
                UNROLL for (int i = 0; i < NDIMS; i++)
                    V[n][i] = VS[n*(NDIMS*C_SUB) + COFF(e, i, NDIMS)];
            UNROLL for (int i = 0; i < NDIMS; i++)
                x[i] = PTS[p*NDIMS + i];
            calc_smats_detj(V, x, s, djac);
            const fpdtype_t rcpdjac_v = FP(1.0)/djac;
'''
            npt_words, q_words = nu*nd, 0

        npt_words += npt_words % 2
        geo_words = 2*nverts*nd*csub + npt_words + q_words
        geo_words += (-geo_words*isz // 16 * -16 - geo_words*isz)//isz
        geo_decl = f'''fpdtype_t *VSB = G + G_WORDS;
    fpdtype_t *PTS = VSB + 2*V_WORDS;
    fpdtype_t *QS = PTS + {npt_words};'''
        geo_stage = f'''for (int i = tid; i < {npt_words}; i += NTHREADS)
        PTS[i] = c_pts[i];'''
        geo_fetch = '''tma_load_1d(VSB + (n & 1)*V_WORDS, verts + b*verts_bsz,
                    V_WORDS*sizeof(fpdtype_t), &bars[0]);'''
        geo_bytes = ' + V_WORDS*sizeof(fpdtype_t)'
        geo_blk = 'const fpdtype_t *VS = VSB + (it & 1)*V_WORDS;'
    else:
        geo_words, geo_decl, geo_stage, geo_fetch = 0, '', '', ''
        geo_bytes, geo_blk = '', ''
        gsrc = ''
        gargs = ('const fpdtype_t* __restrict__ smats, long long smats_bsz, '
                 'const fpdtype_t* __restrict__ rcpdjac, '
                 'long long rcpdjac_bsz')
        geom = r'''
            fpdtype_t s[NDIMS][NDIMS];
            UNROLL for (int i = 0; i < NDIMS; i++)
                UNROLL for (int j = 0; j < NDIMS; j++)
                    s[i][j] = __ldg(smats + blk*smats_bsz
                                    + (long long) (i*NPTS + p)*(NDIMS*C_SUB)
                                    + COFF(e, j, NDIMS));
            const fpdtype_t rcpdjac_v = __ldg(rcpdjac + blk*rcpdjac_bsz
                                              + p*C_SUB + e);
'''

    return dict(gsrc=gsrc, gargs=gargs, geo_elem=geo_elem, geo_post=geo_post,
                geom=geom, geo_words=geo_words, geo_decl=geo_decl,
                geo_stage=geo_stage, geo_fetch=geo_fetch,
                geo_bytes=geo_bytes, geo_blk=geo_blk)


def gradflux_source(be, ops, tplargs, pts, LD, nthreads=None, rowcls=None,
                    affine=False):
    """Source of the fused kernel.

    ``rowcls``: optional class index (< 16) per flux-point row; the kernel
    then takes a per-block bit mask ``fmask`` and neither computes nor
    stores the gradients of rows whose class bit is clear (rows no
    interface kernel ever reads, see fusion.row_need_classes).

    ``affine``: every element of the region has a constant Jacobian (the
    caller has checked the vertices); the metric terms are then formed once
    per block by each thread for the one element it works on and kept in
    registers, instead of being re-evaluated from shared memory at every
    point in phases 2 and 4.

    ``ops``: dict with the operator matrices ``A1`` (ndims*nupts x nupts),
    ``M6`` (ndims*nupts x nfpts), ``M0`` (nfpts x nupts) and ``A5``
    (nupts x ndims*nupts); ``tplargs``: the tflux template arguments
    (``ktype`` is 'linear' or 'curved').  Raises ``NotFusable`` when the
    operators lack the line structure the kernel relies on."""
    nd, nv = tplargs['ndims'], tplargs['nvars']
    A1, M6, M0, A5 = (np.asarray(ops[k], dtype=float)
                      for k in ('A1', 'M6', 'M0', 'A5'))
    nu, nf = M0.shape[1], M0.shape[0]
    isz = np.dtype(be.fpdtype).itemsize
    csub = be.csubsz

    if A1.shape != (nd*nu, nu) or M6.shape != (nd*nu, nf) or \
       A5.shape != (nu, nd*nu) or LD != nv*csub:
        raise NotFusable('unexpected operator shapes')

    linear = 'linear' in tplargs['ktype']
    npoints = nu*csub

    # Occupancy plan: as many CTAs per SM as the shared-memory footprint
    # allows (two when a block is half an SM's worth), sharing a budget of
    # 512 threads so that each keeps ~128 registers.  Co-resident CTAs run
    # different phases at any one time, which overlaps the FP64-heavy
    # pointwise phases of one with the shared-memory/HBM phases of another.
    smem_fix = ((nu + nf + nd*nu)*LD + (2*tplargs.get('nverts', 0)*nd*csub
                                        + 2*nu*nd + 16*nd*csub if linear else 0))*isz + 64
    smem_sm = 228*1024
    nctas = max(1, min(getattr(be, 'gradflux_maxctas', 2),
                       smem_sm // (smem_fix + 1024)))
    if nthreads is None:
        nthreads = getattr(be, 'gradflux_threads', 0) or 512 // nctas
    nrounds = -(-npoints // nthreads)

    defs = [('NDIMS', nd), ('NVARS', nv), ('NPTS', nu), ('NFPTS', nf),
            ('NVERTS', tplargs.get('nverts', 0)), ('NEED_RCPDJAC', 1),
            ('LD', LD), ('NTHREADS', nthreads), ('NROUNDS', nrounds)]
    defs += ph.physics_defines(tplargs['c'], tplargs.get('visc_corr', 'none'),
                               True)

    K = ConstPool(isz == 8)
    em = PhaseEmitter(LD, isz, K, ncol=getattr(be, 'gradflux_ncol', 1))

    vecs = set(getattr(be, 'gradflux_vec2', ()) or ())

    # Phase 1: G = A1 @ U + M6 @ C
    p1 = em.emit('p1', build_classes([A1, M6]), ['U', 'C'],
                 lambda off, v, at: f'{at("G", off)} = {v};',
                 vec='p1' in vecs)

    # Phase 3: vect_fpts[d] = M0 @ G[d]
    M0d = np.zeros((nd*nf, nd*nu))
    for d in range(nd):
        M0d[d*nf:(d + 1)*nf, d*nu:(d + 1)*nu] = M0
    if rowcls is not None and max(rowcls) < 16:
        outtag = lambda r: int(rowcls[r % nf])
        fm_arg = ',\n         const int* __restrict__ fmask'
        fm_load = 'const unsigned fm = (unsigned) __ldg(fmask + blk);'
    else:
        outtag, fm_arg, fm_load = None, '', ''

    p3 = em.emit('p3', build_classes([M0]), ['G'],
                 lambda off, v, at: f'{at("(vf + vfb)", off)} = {v};',
                 outtag=outtag, rep=(nd, nf*LD*isz, nu*LD*isz),
                 vec='p3' in vecs)

    # Phase 5: in-place line transforms of the flux, direction by
    # direction (block d of A5 acts on rows d*nu.. of G), then the sum
    A5d = np.zeros((nd*nu, nd*nu))
    for d in range(nd):
        A5d[d*nu:(d + 1)*nu, d*nu:(d + 1)*nu] = A5[:, d*nu:(d + 1)*nu]
    p5lines = em.emit('p5', build_classes([A5d]), ['G'],
                      lambda off, v, at: f'{at("G", off)} = {v};',
                      inplace=True, vec='p5' in vecs)
    if LD % 2 == 0:
        gv = 'reinterpret_cast<const fpdtype2_t *>(G)'
        sx = ' + '.join(f'{gv}[{d*nu*LD // 2} + item].x' for d in range(nd))
        sy = ' + '.join(f'{gv}[{d*nu*LD // 2} + item].y' for d in range(nd))
        p5 = f'''{p5lines}
        __syncthreads();

        for (int item = tid; item < NPTS*LD/2; item += NTHREADS)
        {{
            fpdtype2_t t;
            t.x = {sx};
            t.y = {sy};
            reinterpret_cast<fpdtype2_t *>(fout + fob)[item] = t;
        }}'''
    else:
        psum = ' + '.join(f'G[{d*nu*LD} + item]' for d in range(nd))
        p5 = f'''{p5lines}
        __syncthreads();

        for (int item = tid; item < NPTS*LD; item += NTHREADS)
            fout[fob + item] = {psum};'''

    # Preferred form: the last direction by in-place line transforms, the
    # first two accumulated plane by plane in registers and streamed out
    if getattr(be, 'gradflux_planes', True):
        blocks = [A5[:, d*nu:(d + 1)*nu] for d in range(nd)]
        try:
            mark = len(em.tables)
            if nd == 3:
                A5l = np.zeros_like(A5d)
                A5l[2*nu:, 2*nu:] = blocks[2]
                # Rows of the other directions are left untouched
                lines = build_classes([A5l], rows=range(2*nu, 3*nu))
                p5a = em.emit('p5l', lines, ['G'],
                              lambda off, v, at: f'{at("G", off)} = {v};',
                              inplace=True, vec='p5' in vecs
                              ) + '\n        __syncthreads();'
                extra = lambda off, at: at('G', f'{off} + {2*nu*LD*isz}')
            else:
                p5a, extra = '', None

            p5b = em.emit_planes(
                'p5p', blocks[:2], [0, nu], extra,
                lambda off, v, at: f'{at("(fout + fob)", off)} = {v};'
            )
            # Drop the tables of the line-only variant
            em.tables = [t for t in em.tables[:mark]
                         if not t[0].startswith('tab_p5_')] + em.tables[mark:]
            p5 = f'{p5a}\n        {p5b}'
        except NotFusable:
            em.tables = em.tables[:mark]

    geo = geometry_source(be, tplargs, pts, nthreads, affine)
    gsrc, gargs, geo_elem, geo_post = (geo[k] for k in (
        'gsrc', 'gargs', 'geo_elem', 'geo_post'))
    geom, geo_words, geo_decl, geo_stage = (geo[k] for k in (
        'geom', 'geo_words', 'geo_decl', 'geo_stage'))
    geo_fetch, geo_bytes, geo_blk = (geo[k] for k in (
        'geo_fetch', 'geo_bytes', 'geo_blk'))

    # Index tables go to shared memory as far as the per-CTA share of the
    # SM's 228 KB (1 KB of it reserved per resident CTA) allows
    tables = em.decls()
    data_words = (nu + nf + nd*nu) + 0
    smem = (data_words*LD + geo_words)*isz + 64
    em.plan(min(smem_sm // nctas - 1024, 227*1024) - smem)
    smem += em.table_bytes

    if smem > 227*1024:
        raise NotFusable(f'needs {smem} bytes of shared memory')

    src = f'''{ph.prologue(be.fpdtype.__name__, be.ixdtype.__name__,
                          be.soasz, be.csubsz, defs)}
{_pipeline_src}
{ph.flux_src}
{ph.visc_src}
{ph.geom_src}
{gsrc}
{tables}
{K.decl()}

#define U_WORDS (NPTS*LD)
#define C_WORDS (NFPTS*LD)
#define G_WORDS (NDIMS*NPTS*LD)
#define V_WORDS (NVERTS*NDIMS*C_SUB)

extern "C" __global__ void __launch_bounds__(NTHREADS, {nctas})
gradflux(int nblocks, int neles,
         const fpdtype_t* __restrict__ u, long long u_bsz,
         const fpdtype_t* ucomm, long long ucomm_bsz,
         fpdtype_t* vf, long long vf_bsz,
         fpdtype_t* __restrict__ fout, long long fout_bsz,
         {gargs}{fm_arg})
{{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    fpdtype_t *U = reinterpret_cast<fpdtype_t *>(smem_raw);
    fpdtype_t *C = U + U_WORDS;
    fpdtype_t *G = C + C_WORDS;
    {geo_decl}
    {em.smem_layout(f'G + G_WORDS + {geo_words}')}
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(
        reinterpret_cast<unsigned char *>(G + G_WORDS + {geo_words})
        + {em.table_bytes});

    const int tid = threadIdx.x;

    // Stage the index tables and the reference point set once per CTA
    {em.stage(f'G + G_WORDS + {geo_words}')}
    {geo_stage}

    if (tid == 0)
    {{
        mbar_init(&bars[0], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }}
    __syncthreads();

    auto fetch = [&](long long b, unsigned n)
    {{
        mbar_expect_tx(&bars[0], (U_WORDS + C_WORDS)*sizeof(fpdtype_t)
                                 {geo_bytes});
        tma_load_1d(U, u + b*u_bsz, U_WORDS*sizeof(fpdtype_t), &bars[0]);
        tma_load_1d(C, ucomm + b*ucomm_bsz, C_WORDS*sizeof(fpdtype_t),
                    &bars[0]);
        {geo_fetch}
    }};

    long long blk = blockIdx.x;
    if (tid == 0 && blk < nblocks)
        fetch(blk, 0);

    for (unsigned it = 0; blk < nblocks; blk += gridDim.x, it++)
    {{
        const long long nxt = blk + gridDim.x;
        const long long vfb = blk*vf_bsz, fob = blk*fout_bsz;

        mbar_wait(&bars[0], it & 1);
        {geo_blk}
        {fm_load}
{geo_elem}

        // ---- phase 1: corrected transformed gradient ------------------
        {p1}

        // Keep the solution at this thread's flux-evaluation points
        fpdtype_t ureg[NROUNDS][NVARS];
        UNROLL for (int r = 0; r < NROUNDS; r++)
        {{
            const int item = tid + r*NTHREADS;
            if (item < NPTS*C_SUB)
            {{
                const int e = item % C_SUB, p = item / C_SUB;
                UNROLL for (int v = 0; v < NVARS; v++)
                    ureg[r][v] = U[p*LD + COFF(e, v, NVARS)];
            }}
        }}
        __syncthreads();
{geo_post}
        // u and ucomm are consumed: fetch the next block's behind the
        // remaining phases
        if (tid == 0 && nxt < nblocks)
            fetch(nxt, it + 1);

        // ---- phase 2: physical gradient (in place) ---------------------
        for (int item = tid; item < NPTS*C_SUB; item += NTHREADS)
        {{
            const int e = item % C_SUB, p = item / C_SUB;
            if (blk*C_SUB + e >= neles)
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
        __syncthreads();

        // ---- phase 3: gradients at the flux points -> HBM ---------------
        {p3}
        __syncthreads();

        // ---- phase 4: transformed flux (in place over the gradient) -----
        UNROLL for (int r = 0; r < NROUNDS; r++)
        {{
            const int item = tid + r*NTHREADS;
            const int e = item % C_SUB, p = item / C_SUB;
            if (item < NPTS*C_SUB && blk*C_SUB + e < neles)
            {{
{geom}
                (void) rcpdjac_v;
                fpdtype_t g[NDIMS][NVARS];
                UNROLL for (int d = 0; d < NDIMS; d++)
                    UNROLL for (int v = 0; v < NVARS; v++)
                        g[d][v] = G[(d*NPTS + p)*LD + COFF(e, v, NVARS)];

                fpdtype_t ft[NDIMS][NVARS], fo[NDIMS][NVARS], pr, vel[NDIMS];
                inviscid_flux(ureg[r], ft, pr, vel);
                viscous_flux_add(ureg[r], g, ft);
                transform_flux(ft, s, fo);

                UNROLL for (int d = 0; d < NDIMS; d++)
                    UNROLL for (int v = 0; v < NVARS; v++)
                        G[(d*NPTS + p)*LD + COFF(e, v, NVARS)] = fo[d][v];
            }}
        }}
        __syncthreads();

        // ---- phase 5: divergence: line transforms, then the sum -> HBM ---
        {p5}
        __syncthreads();
    }}
}}
'''
    meta = dict(nthreads=nthreads, smem=smem, nctas=nctas,
                words_per_block=(2*nu + nf + nd*nf)*LD)

    return src, 'gradflux', meta
