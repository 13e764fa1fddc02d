"""Property tests of the storage layout and view index arithmetic
("int -- must be bit-exact"): the contract
mirror ``pyfr_b200.base`` against the reference's own
``pyfr.backends.base`` on random shapes, SoA widths, block sizes and view
maps.  Both are instantiated over the NumPy oracle storage, so raw bytes,
``get()`` round trips and the ``mapping`` / ``rstrides`` arrays handed to
kernels can be compared directly.  Needs /root/reference."""

import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import refharness as rh
from oracle.npbackend import make_backend
from pyfr_b200 import base
from pyfr_b200.host.config import Config

pytestmark = pytest.mark.skipif(not rh.available(),
                                reason='needs /root/reference')


def _backends(soasz, csubsz, blocks):
    rh.install_stubs()
    import pyfr.backends.base as rbase
    from pyfr.inifile import Inifile

    txt = (f'[backend]\nprecision = double\n[backend-oracle]\n'
           f'soasz = {soasz}\ncsubsz = {csubsz}\nblocks = {int(blocks)}\n')
    return (make_backend(rbase, name='oracle-ref')(Inifile(txt)),
            make_backend(base)(Config(txt)))


layout = st.tuples(st.sampled_from([2, 4, 8, 16]), st.integers(1, 4),
                   st.booleans())


@settings(max_examples=60, deadline=None,
          suppress_health_check=[HealthCheck.too_slow])
@given(layout=layout, nrow=st.integers(1, 9), nvars=st.integers(1, 5),
       neles=st.integers(1, 70), seed=st.integers(0, 2**31))
def test_matrix_layout_matches_reference(layout, nrow, nvars, neles, seed):
    soasz, mult, blocks = layout
    rbe, mbe = _backends(soasz, soasz*mult, blocks)
    ary = np.random.default_rng(seed).standard_normal((nrow, nvars, neles))

    mats = []
    for be in (rbe, mbe):
        m = be.matrix((nrow, nvars, neles), ary, tags={'align'})
        x = be.xchg_matrix((nvars, neles), ary[0])
        c = be.const_matrix(ary[:, 0], tags={'align'})
        be.commit()
        mats.append((m, x, c))

    for r, m in zip(*mats):
        for a in ('nrow', 'ncol', 'leaddim', 'nblocks', 'blocksz', 'nbytes',
                  'datashape', 'ioshape', 'itemsize'):
            assert getattr(r, a) == getattr(m, a), a
        assert tuple(r.traits) == tuple(m.traits)
        assert np.array_equal(r.data, m.data)            # raw storage image
        assert np.array_equal(r.get(), m.get())

    # row slices address the same storage
    (rm, *_), (mm, *_) = mats
    ra, rb = sorted(np.random.default_rng(seed + 1).integers(0, nrow + 1, 2))
    if rb > ra:
        rs, ms = rm.slice(ra, rb), mm.slice(ra, rb)
        assert (rs.offset, rs.nrow, rs.ncol) == (ms.offset, ms.nrow, ms.ncol)


@settings(max_examples=60, deadline=None,
          suppress_health_check=[HealthCheck.too_slow])
@given(layout=layout, nrow=st.integers(2, 9), nvars=st.integers(1, 5),
       neles=st.integers(1, 70), n=st.integers(1, 40), nvrow=st.integers(1, 3),
       seed=st.integers(0, 2**31))
def test_view_indices_match_reference(layout, nrow, nvars, neles, n, nvrow,
                                      seed):
    """``View.mapping`` / ``rstrides`` (pyfr/backends/base/types.py:
    294-320) and the packed layout of an exchange view, over two matrices
    in one extent."""
    soasz, mult, blocks = layout
    rbe, mbe = _backends(soasz, soasz*mult, blocks)
    rng = np.random.default_rng(seed)

    nvrow = min(nvrow, nrow)
    which = rng.integers(0, 2, n)
    rmap = rng.integers(0, nrow - nvrow + 1, n)
    cmap = rng.integers(0, neles, n)
    # row strides that keep every view row inside the matrix
    rmax = np.maximum((nrow - 1 - rmap)//max(nvrow - 1, 1), 1)
    rstri = 1 + rng.integers(0, 1 << 30, n) % rmax

    out = []
    for be in (rbe, mbe):
        ms = [be.matrix((nrow, nvars, neles), extent='shared',
                        tags={'align'}) for _ in range(2)]
        be.commit()
        matmap = np.array([ms[w].mid for w in which])

        vshape = (nvrow, nvars) if nvrow > 1 else (nvars,)
        kw = dict(rstridemap=rstri) if nvrow > 1 else {}
        v = be.view(matmap, rmap, cmap, vshape=vshape, **kw)
        xv = be.xchg_view(matmap, rmap, cmap, vshape=vshape, **kw)
        be.commit()

        # matrix ids differ between the two backends: compare offsets
        # relative to the first matrix of the extent
        out.append((v.mapping.get() - ms[0].offset//ms[0].itemsize,
                    v.rstrides.get() if nvrow > 1 else None,
                    (xv.xchgmat.nrow, xv.xchgmat.ncol, xv.xchgmat.leaddim),
                    (v.n, v.nvrow, v.nvcol)))

    (rmapg, rstr, rx, rv), (mmapg, mstr, mx, mv) = out
    assert np.array_equal(rmapg, mmapg)
    assert (rstr is None and mstr is None) or np.array_equal(rstr, mstr)
    assert rx == mx and rv == mv


@settings(max_examples=80, deadline=None)
@given(n=st.integers(1, 60), ndims=st.integers(1, 3),
       ngrid=st.integers(1, 6), seed=st.integers(0, 2**31))
def test_fuzzy_sort_and_clean_match_reference(n, ndims, ngrid, seed):
    """``fuzzy_lexsort`` (flux-point ordering agreed by both sides of an
    interface) against ``pyfr.nputil.fuzzysort`` on point clouds with many
    nearly-coincident coordinates, and ``clean`` (constants snapped before
    they are baked into kernels) against ``pyfr.nputil.clean``."""
    rh.install_stubs()
    from pyfr.nputil import batched_fuzzysort, clean as rclean

    from pyfr_b200.host.elements import fuzzy_lexsort
    from pyfr_b200.host.shapes import clean

    rng = np.random.default_rng(seed)

    # points on a coarse lattice plus round-off sized noise: ties in the
    # leading coordinates must be broken by the later ones
    pts = rng.integers(0, ngrid, (n, ndims)).astype(float)
    pts += 1e-12*rng.standard_normal(pts.shape)
    _, ix = np.unique(np.round(pts, 6), axis=0, return_index=True)
    pts = pts[np.sort(ix)]                      # distinct lattice sites

    # (neles, ndims, npts): the same cloud seen by three elements, two of
    # them with their own round-off
    coords = np.stack([pts.T, pts.T + 1e-13, pts.T[:, ::-1]])
    assert np.array_equal(fuzzy_lexsort(coords), batched_fuzzysort(coords))

    a = rng.choice([0.0, 1e-13, 0.5, -0.5, 0.5 + 1e-12, 1/3, -1/3 + 2e-13,
                    2.0, rng.standard_normal()], size=(n, 3))
    assert np.array_equal(clean(a), rclean(lambda: a)())


@settings(max_examples=120, deadline=None)
@given(nk=st.integers(1, 14), seed=st.integers(0, 2**31))
def test_graph_schedule_matches_reference(nk, seed):
    """Row a20: the kernel order a graph commits to -- dependency DAG,
    kernels an exchange send waits on scheduled first, grouped kernels kept
    together (pyfr/backends/base/types.py:343-533) -- on random graphs:
    random hard / pseudo dependencies, sends hanging off random kernels,
    receives, random contiguous groups."""
    rh.install_stubs()
    import pyfr.backends.base as rbase

    rng = np.random.default_rng(seed)
    deps = [sorted(rng.choice(i, rng.integers(0, min(i, 3) + 1),
                              replace=False).tolist()) if i else []
            for i in range(nk)]
    pdeps = [sorted(rng.choice(i, rng.integers(0, 2), replace=False).tolist())
             if i else [] for i in range(nk)]
    sends = [sorted(rng.choice(nk, min(nk, int(rng.integers(1, 3))),
                               replace=False).tolist())
             for _ in range(rng.integers(0, 3))]
    nrecv = int(rng.integers(0, 3))

    # groups: chains k -> k+1 (-> k+2) where each member depends on the one
    # before it, as the solvers hand them over
    groups, used, i = [], set(), 0
    while i < nk - 1:
        glen = int(rng.integers(1, 4))
        g = list(range(i, min(i + glen, nk)))
        if len(g) > 1 and rng.random() < 0.5:
            for a, b in zip(g, g[1:]):
                if a not in deps[b]:
                    deps[b] = sorted(deps[b] + [a])
            groups.append(g)
        i += glen

    programs = []
    for b in (rbase, base):
        be = make_backend(b)(Config('[backend]\nprecision = double\n')
                             if b is base else
                             __import__('pyfr.inifile').inifile.Inifile(
                                 '[backend]\nprecision = double\n'))
        ks = [be.kernel_cls(lambda: None) for _ in range(nk)]
        label = {id(k): f'k{i}' for i, k in enumerate(ks)}

        class Req:
            def __init__(self, name):
                self.name = name

        g = be.graph()
        g.add_mpi_reqs([Req(f'r{j}') for j in range(nrecv)])
        for i, k in enumerate(ks):
            g.add(k, deps=[ks[d] for d in deps[i]],
                  pdeps=[ks[d] for d in pdeps[i]])
        for j, sd in enumerate(sends):
            g.add_mpi_req(Req(f's{j}'), deps=[ks[d] for d in sd])
        for grp in groups:
            g.group([ks[i] for i in grp])
        g.commit()

        programs.append([label[id(o)] if w == 'kernel' else o.name
                         for w, o in g.program])

    assert programs[0] == programs[1]
