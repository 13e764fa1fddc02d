"""The oracle's arithmetic against the reference's own kernel templates.

``oracle/minimako.py`` renders the reference's ``.mako`` kernel files
(read from /root/reference, helper functions from the reference's
``makoutil``); the resulting C -- whole kernel bodies with every macro
expanded -- is compiled with gcc and evaluated on random states, and
``oracle/physics.py`` must reproduce it to round-off.  This pins the
restated flux, Riemann-solver, LDG, boundary-state and geometry arithmetic
on the reference itself rather than on a reading of it."""

import ctypes as ct
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import physics as ph
from oracle import refharness as rh

pytestmark = pytest.mark.skipif(not rh.available(),
                                reason='needs /root/reference')

CONSTS = {'gamma': 1.4, 'mu': 3e-3, 'Pr': 0.71, 'cpTref': 2.5, 'cpTs': 1.1}
RTOL = 2e-13


def _spec(text):
    """'inout view fpdtype_t[3][5]' -> (intent, dims)"""
    intent = text.split()[0]
    dims = tuple(int(d) for d in re.findall(r'\[(\d+)\]', text))
    return intent, dims


class CLib:
    """Kernel bodies / macro expansions compiled as C functions taking an
    array of pointers, one per argument."""

    def __init__(self):
        self.fns, self.src = {}, ['#include <math.h>\n#include <string.h>\n'
                                  'typedef double fpdtype_t;\n'
                                  'typedef int ixdtype_t;\n'
                                  # (as the reference's C backends do)
                                  '#define min(a, b) ((a) < (b) ? (a) : (b))\n'
                                  '#define max(a, b) ((a) > (b) ? (a) : (b))\n']

    def add(self, name, args, body):
        """``args``: {arg: (intent, dims)} in call order."""
        pre, post = [], []
        for i, (a, (intent, dims)) in enumerate(args.items()):
            if dims:
                shp = ''.join(f'[{d}]' for d in dims)
                pre.append(f'fpdtype_t {a}{shp}; memcpy({a}, p[{i}], '
                           f'sizeof({a}));')
                if 'out' in intent:
                    post.append(f'memcpy(p[{i}], {a}, sizeof({a}));')
            else:
                pre.append(f'fpdtype_t {a} = *p[{i}];')
                if 'out' in intent:
                    post.append(f'*p[{i}] = {a};')

        self.src.append(f'void {name}(double **p)\n{{\n' + '\n'.join(pre) +
                        f'\n{body}\n' + '\n'.join(post) + '\n}\n')
        self.fns[name] = args

    def build(self):
        d = tempfile.mkdtemp(prefix='pyfr_b200_tpl_')
        c, so = os.path.join(d, 'k.c'), os.path.join(d, 'k.so')
        with open(c, 'w') as f:
            f.write('\n'.join(self.src))
        res = subprocess.run(['gcc', '-O0', '-ffp-contract=off', '-w',
                              '-shared', '-fPIC', '-o', so, c, '-lm'],
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[:3000]
        self.lib = ct.CDLL(so)
        return self

    def call(self, name, **vals):
        """Returns {arg: array} after the call (inputs copied)."""
        args = self.fns[name]
        bufs = [np.array(vals[a], dtype=float).reshape(dims or (1,)).copy()
                for a, (_, dims) in args.items()]
        ptrs = (ct.POINTER(ct.c_double)*len(bufs))(
            *[b.ctypes.data_as(ct.POINTER(ct.c_double)) for b in bufs])
        getattr(self.lib, name)(ptrs)
        return {a: (b if dims else b[0])
                for (a, (_, dims)), b in zip(args.items(), bufs)}


def _kernel(lib, fname, mod, kname, tplargs, extrns=()):
    from oracle.minimako import Renderer

    r = Renderer(tplargs, extrns)
    r.include(mod)
    attrs, body = r.kernels[kname]
    lib.add(fname, {a: _spec(v) for a, v in attrs.items()}, body)
    return r


def _state(rng, nd, mach=0.5):
    rho = 1 + 0.3*rng.random()
    v = mach*rng.standard_normal(nd)
    p = 1 + 0.3*rng.random()
    return np.array([rho, *(rho*v), p/(CONSTS['gamma'] - 1)
                     + 0.5*rho*(v @ v)])


def _cols(a):
    """Per-variable 1-element arrays, as the oracle's physics expects."""
    return [np.array([x]) for x in a]


def _close(got, want, scale=None):
    got, want = np.asarray(got, float), np.asarray(want, float)
    scale = scale or max(np.abs(want).max(), 1.0)
    assert np.abs(got - want).max() <= RTOL*scale, (got, want)


# -- interior interface kernels ------------------------------------------------
CFLUX = [(nd, rs, beta, tau, vc) for nd in (2, 3)
         for rs, beta, tau, vc in [('rusanov', 0.5, 0.1, 'none'),
                                   ('hllc', 0.0, 0.0, 'none'),
                                   ('rusanov', -0.5, 0.3, 'sutherland'),
                                   ('hllc', 0.25, 0.1, 'sutherland')]]


@pytest.fixture(scope='module')
def cflux_lib():
    lib = CLib()
    for i, (nd, rs, beta, tau, vc) in enumerate(CFLUX):
        c = dict(CONSTS, **{'ldg-beta': beta, 'ldg-tau': tau})
        tpl = dict(ndims=nd, nvars=nd + 2, c=c, rsolver=rs, visc_corr=vc,
                   shock_capturing='none')
        _kernel(lib, f'ns_intcflux_{i}',
                'pyfr.solvers.navstokes.kernels.intcflux', 'intcflux', tpl)
        _kernel(lib, f'ns_intconu_{i}',
                'pyfr.solvers.navstokes.kernels.intconu', 'intconu', tpl)
        _kernel(lib, f'eu_intcflux_{i}',
                'pyfr.solvers.euler.kernels.intcflux', 'intcflux', tpl)
    return lib.build()


@pytest.mark.parametrize('i', range(len(CFLUX)))
def test_interface_kernels_match_reference_templates(cflux_lib, i):
    nd, rs, beta, tau, vc = CFLUX[i]
    nv = nd + 2
    c = dict(CONSTS, **{'ldg-beta': beta, 'ldg-tau': tau})
    rng = np.random.default_rng(100 + i)

    for _ in range(25):
        ul, ur = _state(rng, nd), _state(rng, nd)
        gl, gr = rng.standard_normal((2, nd, nv))
        nl = rng.standard_normal(nd)

        # Navier-Stokes common flux (Riemann solve + LDG viscous flux)
        out = cflux_lib.call(f'ns_intcflux_{i}', ul=ul, ur=ur, gradul=gl,
                             gradur=gr, artvisc=0.0, nl=nl)
        fn = ph.ns_common_flux(_cols(ul), _cols(ur),
                               [_cols(g) for g in gl], [_cols(g) for g in gr],
                               _cols(nl), nd, nv, c, rs, vc)
        want = np.array([f[0] for f in fn])
        _close(out['ul'], want)
        _close(out['ur'], -want)

        # LDG common solution
        # (a side the kernel does not write keeps its marker value)
        out = cflux_lib.call(f'ns_intconu_{i}', ulin=ul, urin=ur,
                             ulout=np.full(nv, 7.0), urout=np.full(nv, 7.0))
        lo, ro = ph.ldg_common_solution(_cols(ul), _cols(ur), beta)
        for got, want in ((out['ulout'], lo), (out['urout'], ro)):
            _close(got, np.full(nv, 7.0) if want is None else
                   [x[0] for x in want])

        # Euler common flux
        out = cflux_lib.call(f'eu_intcflux_{i}', ul=ul, ur=ur, nl=nl)
        fn = ph.euler_common_flux(_cols(ul), _cols(ur), _cols(nl), nd, nv, c,
                                  rs)
        want = np.array([f[0] for f in fn])
        _close(out['ul'], want)
        _close(out['ur'], -want)
