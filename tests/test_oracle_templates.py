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


# -- boundary kernels ------------------------------------------------------------
def _record_bc_tplargs(system, n, bcs, edits=(), **kw):
    """The template arguments the host code hands to the boundary kernels
    (section constants compiled to C expressions included)."""
    from pyfr_b200 import cases
    from pyfr_b200.host.system import get_system
    from util import OracleBackend

    cfg, box, _ = cases.box_case(system, n, bcs, **kw)
    for sect, opt, val in edits:
        cfg.set(sect, opt, val)

    be = OracleBackend(cfg)
    seen, orig = [], be.kernel

    def kernel(name, *a, **k):
        if name in ('bcconu', 'bccflux'):
            seen.append((name, k['tplargs']))
        return orig(name, *a, **k)

    be.kernel = kernel
    get_system(be, box.local_mesh(), cfg, 2)
    return seen


BC_SETS = [
    ('navier-stokes', (2, 2, 2),
     {'xlo': 'sub-in-frv', 'xhi': 'sub-out-fp', 'ylo': 'no-slp-adia-wall',
      'yhi': 'char-riem-inv', 'zlo': 'slp-adia-wall',
      'zhi': 'no-slp-isot-wall'},
     [('soln-bcs-xlo', 'u', '0.2 + 0.1*sin(3*t) + 0.05*y'),
      ('soln-bcs-zhi', 'u', '0.1 + 0.02*x*cos(t)')],
     dict(order=1, rsolver='hllc', beta=0.0)),
    ('navier-stokes', (2, 2, 2),
     {'xlo': 'sub-in-ftpttang', 'xhi': 'sup-out-fn', 'ylo': 'sup-in-fa',
      'yhi': 'sub-out-fp'},
     [('soln-bcs-yhi', 'p', '71.0 + 0.5*cos(t)')],
     dict(order=1)),
    ('euler', (2, 2),
     {'xlo': 'char-riem-inv', 'xhi': 'sup-out-fn', 'ylo': 'slp-adia-wall',
      'yhi': 'sup-in-fa'},
     [('soln-bcs-yhi', 'rho', '1.0 + 0.1*x*t')], dict(order=1)),
]


def _bc_kernels():
    out = []
    for system, n, bcs, edits, kw in BC_SETS:
        for name, tpl in _record_bc_tplargs(system, n, bcs, edits, **kw):
            out.append((system, name, tpl))
    return out


@pytest.fixture(scope='module')
def bc_lib():
    lib, specs = CLib(), _bc_kernels()
    for i, (system, name, tpl) in enumerate(specs):
        nd = tpl['ndims']
        sysmod = 'navstokes' if system == 'navier-stokes' else 'euler'
        from oracle.minimako import Renderer

        r = Renderer(tpl, extrns=('t', 'ploc'))
        r.include(f'pyfr.solvers.{sysmod}.kernels.{name}')
        attrs, body = r.kernels[name]
        args = {a: _spec(v) for a, v in attrs.items()}
        args['ploc'], args['t'] = ('in', (nd,)), ('in', ())
        lib.add(f'bc_{i}', args, body)
    return lib.build(), specs


def test_boundary_kernels_match_reference_templates(bc_lib):
    lib, specs = bc_lib
    types = set()

    for i, (system, name, tpl) in enumerate(specs):
        nd, nv, c = tpl['ndims'], tpl['nvars'], tpl['c']
        viscous = system == 'navier-stokes'
        rng = np.random.default_rng(500 + i)
        types.add(tpl['bctype'])

        for _ in range(12):
            # mean flow through the face: keeps the states away from the
            # sign switches of the characteristic conditions
            ul = _state(rng, nd, mach=0.2)
            gl = rng.standard_normal((nd, nv))
            nl = rng.standard_normal(nd)
            ploc, t = rng.standard_normal(nd), float(rng.random())
            env = {'t': t, 'ploc': _cols(ploc)}

            if name == 'bcconu':
                out = lib.call(f'bc_{i}', ulin=ul, ulout=np.zeros(nv),
                               nlin=nl, ploc=ploc, t=t)
                mag = np.sqrt(nl @ nl)
                want = ph.bc_ldg_state(tpl['bctype'], _cols(ul),
                                       _cols(nl/mag), nd, nv, c, env)
                _close(out['ulout'], [np.ravel(x)[0] for x in want])
            else:
                kw = dict(ul=ul, nl=nl, ploc=ploc, t=t)
                if viscous:
                    kw.update(gradul=gl, artvisc=0.0)
                out = lib.call(f'bc_{i}', **kw)
                want = ph.bc_common_flux(
                    tpl['bctype'], tpl.get('bccfluxstate'), _cols(ul),
                    [_cols(g) for g in gl] if viscous else None, _cols(nl),
                    nd, nv, c, tpl['rsolver'], env, viscous,
                    tpl.get('visc_corr', 'none')
                )
                _close(out['ul'], [np.ravel(x)[0] for x in want])

    assert types >= {'no-slp-adia-wall', 'no-slp-isot-wall', 'slp-adia-wall',
                     'char-riem-inv', 'sup-in-fa', 'sup-out-fn', 'sub-in-frv',
                     'sub-out-fp', 'sub-in-ftpttang'}
