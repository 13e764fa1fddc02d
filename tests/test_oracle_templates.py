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

# Where the reference is present the rendered kernels are compiled and run;
# elsewhere (the GPU box) the same tests replay the committed input/output
# vectors of such a run (tests/golden/kernel_vectors.npz, written by
# ``make_golden.py --kernels``)
HERE = os.path.dirname(os.path.abspath(__file__))
VECTORS = os.path.join(HERE, 'golden', 'kernel_vectors.npz')
LIVE = rh.available()
RECORD = os.environ.get('PYFR_B200_RECORD_KERNELS')

_recorded, _replay, _cursor = {}, None, {}


@pytest.fixture(scope='module', autouse=True)
def _vectors():
    global _replay
    if not LIVE:
        _replay = np.load(VECTORS)
    yield
    if RECORD:
        out = {}
        for name, calls in _recorded.items():
            for kind, idx in (('in', 0), ('out', 1)):
                for a in calls[0][idx]:
                    out[f'{name}|{kind}|{a}'] = np.stack(
                        [np.asarray(c[idx][a], dtype=float) for c in calls])
        np.savez_compressed(RECORD, **out)


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
        if not LIVE:
            return
        pre, post = [], []
        for i, (a, (intent, dims)) in enumerate(args.items()):
            if dims:
                shp = ''.join(f'[{d}]' for d in dims)
                pre.append(f'fpdtype_t {a}{shp}; memcpy({a}, _argv[{i}], '
                           f'sizeof({a}));')
                if 'out' in intent:
                    post.append(f'memcpy(_argv[{i}], {a}, sizeof({a}));')
            else:
                pre.append(f'fpdtype_t {a} = *_argv[{i}];')
                if 'out' in intent:
                    post.append(f'*_argv[{i}] = {a};')

        self.src.append(f'void {name}(double **_argv)\n{{\n' + '\n'.join(pre) +
                        f'\n{body}\n' + '\n'.join(post) + '\n}\n')
        self.fns[name] = args

    def build(self):
        if not LIVE:
            return self
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
        if not LIVE:
            # replay: same call sequence, same (seeded) inputs
            i = _cursor[name] = _cursor.get(name, -1) + 1
            pre = f'{name}|'
            for k in _replay.files:
                if k.startswith(pre + 'in|'):
                    a = k[len(pre) + 3:]
                    want = _replay[k][i]
                    got = np.asarray(vals[a], dtype=float).reshape(want.shape)
                    assert np.allclose(got, want, rtol=1e-14, atol=0), \
                        f'{name}: input {a} differs from the recorded call'
            return {k[len(pre) + 4:]: _replay[k][i] for k in _replay.files
                    if k.startswith(pre + 'out|')}

        args = self.fns[name]
        bufs = [np.array(vals[a], dtype=float).reshape(dims or (1,)).copy()
                for a, (_, dims) in args.items()]
        ptrs = (ct.POINTER(ct.c_double)*len(bufs))(
            *[b.ctypes.data_as(ct.POINTER(ct.c_double)) for b in bufs])
        ins = [b.copy() for b in bufs]
        getattr(self.lib, name)(ptrs)
        res = {a: (b if dims else b[0])
               for (a, (_, dims)), b in zip(args.items(), bufs)}

        if RECORD:
            _recorded.setdefault(name, []).append((
                {a: (b if dims else b[0])
                 for (a, (_, dims)), b in zip(args.items(), ins)},
                {a: v for a, v in res.items() if 'out' in args[a][0]}
            ))

        return res


def _kernel(lib, fname, mod, kname, tplargs, extrns=()):
    if not LIVE:
        return None

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
        if not LIVE:
            break

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


# -- element kernels (through the oracle backend's kernel objects) --------------
ELE = [(nd, ktype, vc) for nd in (2, 3)
       for ktype, vc in [('curved', 'none'), ('linear', 'sutherland'),
                         ('curved-fused', 'sutherland'),
                         ('linear-fused', 'none')]]


def _ref_jac_exprs(nd):
    if not LIVE:
        # (pinned against the reference's by the host fixtures)
        from pyfr_b200.host.shapes import HexShape, QuadShape
    else:
        rh.install_stubs()
        from pyfr.shapes import HexShape, QuadShape
    return (QuadShape if nd == 2 else HexShape).jac_exprs


@pytest.fixture(scope='module')
def ele_lib():
    lib = CLib()
    for i, (nd, ktype, vc) in enumerate(ELE):
        tpl = dict(ndims=nd, nvars=nd + 2, nverts=2**nd, c=dict(CONSTS),
                   jac_exprs=_ref_jac_exprs(nd), ktype=ktype, visc_corr=vc,
                   shock_capturing='none', src_macros=[])
        _kernel(lib, f'ns_tflux_{i}', 'pyfr.solvers.navstokes.kernels.tflux',
                'tflux', tpl)
        if 'fused' not in ktype:
            _kernel(lib, f'eu_tflux_{i}', 'pyfr.solvers.euler.kernels.tflux',
                    'tflux', tpl)
            _kernel(lib, f'gradcoru_{i}',
                    'pyfr.solvers.baseadvecdiff.kernels.gradcoru', 'gradcoru',
                    tpl)
            _kernel(lib, f'wavespeed_{i}',
                    'pyfr.solvers.euler.kernels.wavespeed', 'wavespeed', tpl)
    _kernel(lib, 'negdivconf', 'pyfr.solvers.baseadvec.kernels.negdivconf',
            'negdivconf', dict(ndims=3, nvars=5, src_macros=[], c=CONSTS),
            extrns=('t', 'ploc', 'u'))
    return lib.build()


@pytest.mark.parametrize('i', range(len(ELE)))
def test_element_kernels_match_reference_templates(ele_lib, i):
    """tflux (Navier-Stokes and Euler, stored and vertex-derived metric
    terms, with and without the fused gradient transform), gradcoru and
    wavespeed: the oracle backend's kernels, run on its blocked storage
    over K one-point elements, against the rendered templates point by
    point."""
    from pyfr_b200.host.config import Config
    from util import OracleBackend

    nd, ktype, vc = ELE[i]
    nv, nverts, K = nd + 2, 2**nd, 21
    linear, fused = 'linear' in ktype, 'fused' in ktype
    rng = np.random.default_rng(900 + i)
    jac = _ref_jac_exprs(nd)
    tpl = dict(ndims=nd, nvars=nv, nverts=nverts, c=dict(CONSTS),
               jac_exprs=jac, ktype=ktype, visc_corr=vc,
               shock_capturing='none', src_macros=[])

    # Random well-conditioned inputs
    u = np.stack([_state(rng, nd) for _ in range(K)], axis=-1)[None]
    g = rng.standard_normal((nd, 1, nv, K))
    ref = np.array([[-1.0, 1.0][(n >> d) & 1] for n in range(nverts)
                    for d in range(nd)]).reshape(nverts, nd)
    verts = (ref[:, :, None]*(1 + 0.2*rng.random((1, nd, K)))
             + 0.15*rng.standard_normal((nverts, nd, K)))
    upts = 0.6*rng.uniform(-1, 1, (1, nd))
    smats = (np.eye(nd)[:, None, :, None]
             + 0.3*rng.standard_normal((nd, 1, nd, K)))
    rcp = 1/(1 + rng.random((1, K)))

    be = OracleBackend(Config('[backend]\nprecision = double\n'
                              '[backend-oracle]\nblocks = 1\nsoasz = 4\n'
                              'csubsz = 8\n'))
    be.pointwise.register('pyfr.solvers.navstokes.kernels.tflux')
    mu, mf = be.matrix(u.shape, u), be.matrix((nd, 1, nv, K), g)
    mg = be.matrix((nd, 1, nv, K), g)
    geo = (dict(verts=be.const_matrix(verts), upts=be.const_matrix(upts))
           if linear else
           dict(smats=be.const_matrix(smats), rcpdjac=be.const_matrix(rcp)))
    be.commit()

    kern = be.kernel('tflux', tplargs=tpl, dims=[1, K], u=mu, f=mf,
                     **(dict(gradu=mg) if fused else {}), **geo)
    kern.run()
    fo, go = mf.get(), mg.get()

    for e in range(K):
        args = dict(u=u[0, :, e], f=g[:, 0, :, e], gradu=g[:, 0, :, e],
                    smats=smats[:, 0, :, e], rcpdjac=rcp[0, e],
                    verts=verts[:, :, e], upts=upts[0],
                    artvisc_vtx=0.0)
        out = ele_lib.call(f'ns_tflux_{i}', **args)
        _close(fo[:, 0, :, e], out['f'])
        if fused:
            _close(go[:, 0, :, e], out['gradu'])

    if fused:
        return

    # Euler flux, gradient transform and wave speed on the same inputs
    be.pointwise._mods.clear()
    for mod in ('pyfr.solvers.euler.kernels.tflux',
                'pyfr.solvers.baseadvecdiff.kernels.gradcoru',
                'pyfr.solvers.euler.kernels.wavespeed'):
        be.pointwise.register(mod)

    mf2, mg2 = be.matrix((nd, 1, nv, K)), be.matrix((nd, 1, nv, K), g)
    mw = be.matrix((1, K))
    be.commit()
    sgeo = {k: v for k, v in geo.items() if k != 'rcpdjac'}
    be.kernel('tflux', tplargs=tpl, dims=[1, K], u=mu, f=mf2, **sgeo).run()
    be.kernel('gradcoru', tplargs=tpl, dims=[1, K], gradu=mg2, **geo).run()
    be.kernel('wavespeed', tplargs=tpl, dims=[1, K], u=mu, wspd=mw,
              **geo).run()
    f2, g2, w = mf2.get(), mg2.get(), mw.get()

    for e in range(K):
        args = dict(u=u[0, :, e], f=np.zeros((nd, nv)),
                    gradu=g[:, 0, :, e], smats=smats[:, 0, :, e],
                    rcpdjac=rcp[0, e], verts=verts[:, :, e], upts=upts[0],
                    wspd=0.0)
        _close(f2[:, 0, :, e], ele_lib.call(f'eu_tflux_{i}', **args)['f'])
        _close(g2[:, 0, :, e],
               ele_lib.call(f'gradcoru_{i}', **args)['gradu'])
        _close(w[0, e], ele_lib.call(f'wavespeed_{i}', **args)['wspd'])


# -- inter-partition kernels, time stepping and diagnostics ---------------------
@pytest.fixture(scope='module')
def misc_lib():
    from pyfr_b200.host.integrator import RK45Stepper as S

    lib = CLib()
    for i, (nd, rs, beta, tau, vc) in enumerate(CFLUX):
        c = dict(CONSTS, **{'ldg-beta': beta, 'ldg-tau': tau})
        tpl = dict(ndims=nd, nvars=nd + 2, c=c, rsolver=rs, visc_corr=vc,
                   shock_capturing='none')
        for k in ('mpicflux', 'mpiconu'):
            _kernel(lib, f'ns_{k}_{i}',
                    f'pyfr.solvers.navstokes.kernels.{k}', k, tpl)
        _kernel(lib, f'eu_mpicflux_{i}',
                'pyfr.solvers.euler.kernels.mpicflux', 'mpicflux', tpl)

    e = [b - bh for b, bh in zip(S.b, S.bhat)]
    for stage in range(5):
        for errest in (False, True):
            _kernel(lib, f'rkvdh2_{stage}_{int(errest)}',
                    'pyfr.integrators.explicit.kernels.rkvdh2', 'rkvdh2',
                    dict(a=S.a, b=S.b, e=e, stage=stage, nstages=5, nvars=4,
                         errest=errest))

    _kernel(lib, 'negdivconf', 'pyfr.solvers.baseadvec.kernels.negdivconf',
            'negdivconf', dict(ndims=3, nvars=5, src_macros=[], c=CONSTS),
            extrns=('t', 'ploc', 'u'))

    for nd in (2, 3):
        exprs = ['pri[0]*pri[1] + pri[%d]' % (nd + 1),
                 'grad_pri[1][0]*grad_pri[%d][1] - t*grad_pri[0][%d]'
                 % (nd + 1, nd - 1)]
        _kernel(lib, f'fieldeval_{nd}', 'pyfr.plugins.kernels.fieldeval',
                'fieldeval',
                dict(ndims=nd, nvars=nd + 2, nexprs=2, exprs=exprs,
                     reduceop='sum', c=CONSTS, has_grads=True,
                     use_views=False, has_wts=True, fpdtype_max='1e300',
                     eos_mod='pyfr.solvers.euler.kernels.eos'))
    return lib.build()


def test_mpi_interface_kernels_match_reference_templates(misc_lib):
    """mpicflux / mpiconu: the left-hand side's view of an inter-partition
    interface (the host flips the sign of beta on one of the two ranks)."""
    for i, (nd, rs, beta, tau, vc) in enumerate(CFLUX):
        nv = nd + 2
        c = dict(CONSTS, **{'ldg-beta': beta, 'ldg-tau': tau})
        rng = np.random.default_rng(300 + i)

        for _ in range(10):
            ul, ur = _state(rng, nd), _state(rng, nd)
            gl, gr = rng.standard_normal((2, nd, nv))
            nl = rng.standard_normal(nd)

            out = misc_lib.call(f'ns_mpicflux_{i}', ul=ul, ur=ur, gradul=gl,
                                gradur=gr, artvisc=0.0, nl=nl)
            fn = ph.ns_common_flux(_cols(ul), _cols(ur),
                                   [_cols(g) for g in gl],
                                   [_cols(g) for g in gr], _cols(nl), nd, nv,
                                   c, rs, vc)
            _close(out['ul'], [f[0] for f in fn])

            # mpiconu always writes its side (the oracle kernel spells the
            # three cases out: oracle/npbackend.py _mpiconu)
            out = misc_lib.call(f'ns_mpiconu_{i}', ulin=ul, urin=ur,
                                ulout=np.full(nv, 7.0))
            want = (ul if beta == -0.5 else ur if beta == 0.5 else
                    ur*(0.5 + beta) + ul*(0.5 - beta))
            _close(out['ulout'], want)

            out = misc_lib.call(f'eu_mpicflux_{i}', ul=ul, ur=ur, nl=nl)
            fn = ph.euler_common_flux(_cols(ul), _cols(ur), _cols(nl), nd, nv,
                                      c, rs)
            _close(out['ul'], [f[0] for f in fn])


def test_time_stepping_kernels_match_reference_templates(misc_lib):
    """rkvdh2 (every stage, with and without the error estimate) and
    negdivconf, through the oracle backend's kernel objects."""
    from pyfr_b200.host.config import Config
    from pyfr_b200.host.integrator import RK45Stepper as S
    from util import OracleBackend

    rng = np.random.default_rng(77)
    K, nv, dt = 13, 4, 0.037
    e = [b - bh for b, bh in zip(S.b, S.bhat)]

    be = OracleBackend(Config('[backend]\nprecision = double\n'
                              '[backend-oracle]\nblocks = 1\nsoasz = 4\n'
                              'csubsz = 8\n'))
    be.pointwise.register('pyfr.integrators.explicit.kernels.rkvdh2')
    be.pointwise.register('pyfr.solvers.baseadvec.kernels.negdivconf')

    for stage in range(5):
        for errest in (False, True):
            vals = rng.standard_normal((4, 1, nv, K))
            ms = [be.matrix((1, nv, K), v) for v in vals]
            be.commit()

            names = ('r1', 'r2', 'rold', 'rerr')[:4 if errest else 2]
            k = be.kernel('rkvdh2', tplargs=dict(
                a=S.a, b=S.b, e=e, stage=stage, nstages=5, nvars=nv,
                errest=errest), dims=[1, K], **dict(zip(names, ms)))
            k.bind(dt=dt)
            k.run()
            got = [m.get() for m in ms]

            for el in range(K):
                out = misc_lib.call(
                    f'rkvdh2_{stage}_{int(errest)}', dt=dt,
                    **{n: vals[j, 0, :, el] for j, n in
                       enumerate(('r1', 'r2', 'rold', 'rerr'))})
                for j, n in enumerate(names):
                    _close(got[j][0, :, el], out[n])

    # negdivconf
    d, r = rng.standard_normal((1, 5, K)), 1/(1 + rng.random((1, K)))
    md, mr = be.matrix(d.shape, d), be.const_matrix(r)
    be.commit()
    k = be.kernel('negdivconf', tplargs=dict(ndims=3, nvars=5, src_macros=[],
                                             c=CONSTS),
                  dims=[1, K], tdivtconf=md, rcpdjac=mr)
    k.run()
    for el in range(K):
        out = misc_lib.call('negdivconf', tdivtconf=d[0, :, el],
                            rcpdjac=r[0, el], ploc=np.zeros(3),
                            u=np.zeros(5), t=0.0)
        _close(md.get()[0, :, el], out['tdivtconf'])


@pytest.mark.parametrize('nd', [2, 3])
def test_fieldeval_matches_reference_template(misc_lib, nd):
    """fieldeval with con_to_pri / grad_con_to_pri of the reference's eos
    template: per-point weighted expression values."""
    from pyfr_b200.host.config import Config
    from util import OracleBackend

    nv, K = nd + 2, 9
    rng = np.random.default_rng(40 + nd)
    exprs = ['pri[0]*pri[1] + pri[%d]' % (nd + 1),
             'grad_pri[1][0]*grad_pri[%d][1] - t*grad_pri[0][%d]'
             % (nd + 1, nd - 1)]
    tpl = dict(ndims=nd, nvars=nv, nexprs=2, exprs=exprs, reduceop='sum',
               c=CONSTS, has_grads=True, use_views=False, has_wts=True,
               eos_mod='pyfr.solvers.euler.kernels.eos')

    u = np.stack([_state(rng, nd) for _ in range(K)], axis=-1)[None]
    g = rng.standard_normal((nd, 1, nv, K))
    w = rng.random((1, K))

    be = OracleBackend(Config('[backend]\nprecision = double\n'))
    be.pointwise.register('pyfr.plugins.kernels.fieldeval')
    mu, mg = be.matrix(u.shape, u), be.matrix(g.shape, g)
    mw, mo = be.const_matrix(w), be.matrix((2, K))
    be.commit()
    k = be.kernel('fieldeval', tplargs=tpl, dims=[1, K], u=mu, gradu=mg,
                  wts=mw, out=mo)
    k.bind(t=0.3)
    k.run()

    for el in range(K):
        out = misc_lib.call(f'fieldeval_{nd}', u=u[0, :, el],
                            gradu=g[:, 0, :, el], ploc=np.zeros(nd),
                            wts=w[0, el], out=np.zeros(2), t=0.3)
        _close(mo.get()[:, el], out['out'])
