"""CPU tests of the kernel generators' host-side logic: the structure they
discover in the operator matrices must reproduce the operators exactly,
and the geometry rewrites must agree with the expressions they replace.
(What the generated CUDA computes is covered by the -m gpu parity tests;
these pin the decompositions the source is rendered from.)"""

import numpy as np
import pytest

from pyfr_b200.host.config import Config
from pyfr_b200.host.shapes import HexShape, QuadShape, shape_map
from pyfr_b200.kernels import fused, mul
from pyfr_b200.kernels import physics as ph


def _shape(et, order, pts='gauss-legendre'):
    face = 'line' if et == 'quad' else 'quad'
    cfg = Config(f'[solver]\norder = {order}\n'
                 f'[solver-elements-{et}]\nsoln-pts = {pts}\n'
                 f'[solver-interfaces-{face}]\nflux-pts = {pts}\n')
    return shape_map[et](None, cfg)


def _apply_classes(classes, srcs, nout):
    """What the emitted class loops compute, in NumPy."""
    out = np.zeros((nout, srcs[0].shape[1]))
    hit = np.zeros(nout, dtype=int)

    for c in classes:
        for rows, ins in c.members:
            acc = sum(c.coefs[t] @ srcs[t][ins[t]] for t in range(len(ins)))
            out[rows] = acc
            hit[rows] += 1

    assert np.all(hit == 1)          # every output row exactly once
    return out


@pytest.mark.parametrize('et,order', [('hex', 2), ('hex', 3), ('hex', 4),
                                      ('quad', 3), ('quad', 5)])
def test_line_classes_reproduce_operators(et, order):
    sh = _shape(et, order)
    nd = sh.ndims
    rng = np.random.default_rng(order)

    M0, M6 = sh.opmat('M0'), sh.opmat('M6')
    A1, A5 = sh.opmat('M4 - M6*M0'), sh.opmat('M1 - M3*M2')
    nu, nf = sh.nupts, sh.nfpts
    u, c = rng.standard_normal((nu, 7)), rng.standard_normal((nf, 7))

    # phase 1: two-term operator
    cls = fused.build_classes([A1, M6])
    assert np.allclose(_apply_classes(cls, [u, c], nd*nu), A1 @ u + M6 @ c,
                       rtol=0, atol=1e-12)
    # few classes, short lines: the point of the decomposition
    assert len(cls) <= 6 and max(sum(k.nins) for k in cls) <= 2*(order + 2)

    # phase 3: interpolation to the flux points
    cls = fused.build_classes([M0])
    assert np.allclose(_apply_classes(cls, [u], nf), M0 @ u, atol=1e-12)

    # phase 5: block-diagonal divergence, lines stay inside their inputs
    A5d = np.zeros((nd*nu, nd*nu))
    for d in range(nd):
        A5d[d*nu:(d + 1)*nu, d*nu:(d + 1)*nu] = A5[:, d*nu:(d + 1)*nu]
    g = rng.standard_normal((nd*nu, 7))
    cls = fused.build_classes([A5d])
    t = _apply_classes(cls, [g], nd*nu)
    assert all(set(r) <= set(i[0]) for k in cls for r, i in k.members)
    assert np.allclose(sum(t[d*nu:(d + 1)*nu] for d in range(nd)), A5 @ g,
                       atol=1e-11)


def test_planes_partition_the_points():
    sh = _shape('hex', 4)
    nu, A5 = sh.nupts, sh.opmat('M1 - M3*M2')
    blocks = [A5[:, d*nu:(d + 1)*nu] for d in range(3)]

    planes = fused.find_planes(blocks[:2])
    assert sorted(p for P in planes for p in P) == list(range(nu))
    assert [len(P) for P in planes] == [25]*5
    # all three directions couple the whole element: too big for a thread
    assert fused.find_planes(blocks) is None


@pytest.mark.parametrize('cls', [HexShape, QuadShape])
def test_monomial_jacobian_equals_expressions(cls):
    nd, nverts = cls.ndims, 2**cls.ndims
    monos, W = ph.multilinear_jacobian(cls.jac_exprs, nd, nverts)
    rng = np.random.default_rng(3)

    for _ in range(5):
        V = rng.standard_normal((nverts, nd))
        x = rng.uniform(-1, 1, nd)
        for d in range(nd):
            for i in range(nd):
                ref = eval(cls.jac_exprs[d][i], {'V': V, 'x': x})
                val = sum(np.prod(x[list(m)])*(W[d, k] @ V[:, i])
                          for k, m in enumerate(monos))
                assert abs(val - ref) < 1e-13


def test_non_multilinear_expressions_are_refused():
    bad = [['x[0]*x[0]*V[0][0]', 'V[1][1]'], ['V[0][0]', 'V[1][1]']]
    assert ph.multilinear_jacobian(bad, 2, 4) is None


def test_affine_detection():
    from pyfr_b200 import cases
    from pyfr_b200.fusion import region_is_affine

    class FakeVerts:
        def __init__(self, v):
            self._v = v

        def get(self):
            return self._v

    for warp, expect in ((0.0, True), (0.05, False)):
        box = cases.tgv_mesh((3, 2, 2), warp=warp)
        v = box.vertices(np.arange(box.neles))          # (8, ne, 3)
        assert region_is_affine(FakeVerts(v.swapaxes(1, 2))) == expect

    # a sheared (still affine) box
    box = cases.tgv_mesh((2, 2, 2))
    v = box.vertices(np.arange(box.neles)).copy()
    v[..., 0] += 0.3*v[..., 1] - 0.2*v[..., 2]
    assert region_is_affine(FakeVerts(v.swapaxes(1, 2)))


def test_chunk_plan_respects_budget():
    for K, LD, isz, budget, hint in [(375, 40, 8, 200*1024, 125),
                                     (1029, 80, 4, 200*1024, 343),
                                     (125, 40, 8, 200*1024, None)]:
        chunks = mul.plan_chunks(K, LD, isz, budget, hint)
        rows = chunks[0][1] - chunks[0][0]

        assert chunks[0][0] == 0 and chunks[-1][1] == K
        assert all(b - a == rows for a, b in chunks)
        assert 2*rows*LD*isz <= budget
