"""CPU tests that execute the product's *generated CUDA kernels* (the
same source text that is compiled for sm_100a) on the host through the
execution model in tests/cudaemu: B200Backend end to end -- host code,
generators, fusion decisions, launch arguments, kernel bodies -- against
the oracle.  Complements the -m gpu parity tests (which run the same
kernels on the device through the C ABI): logic is checked here on every
CPU run, the device-specific behaviour (TMA, occupancy, timing) there."""

import ctypes as ct
import os
import sys

import numpy as np
import pytest

from pyfr_b200 import cases
from pyfr_b200.host.system import get_system

from util import (OracleBackend, assert_parity, oracle_rhs, rel_err,
                  rhs_magnitude)

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                'cudaemu'))
import emu                                                   # noqa: E402


@pytest.fixture
def emulated(monkeypatch):
    emu.install(monkeypatch)


def _b200(cfg, box, vparts=None, rank=0, comm=None, nregs=2, opts={}):
    from pyfr_b200.backend import B200Backend

    cfg.set('backend-b200', 'graphs', os.environ.get('EMU_GRAPHS', 'false'))
    # EMU_OPTS="conu-pairs=1;inters-order=address" runs the whole file with
    # those backend options (as EMU_GRAPHS=true does for graph capture)
    for kv in filter(None, os.environ.get('EMU_OPTS', '').split(';')):
        cfg.set('backend-b200', *kv.split('=', 1))
    for k, v in opts.items():
        cfg.set('backend-b200', k, v)

    be = B200Backend(cfg, comm=comm)
    assert getattr(be.rt, 'emulated', False)
    if comm is not None:
        comm.rt = be.rt
    return get_system(be, box.local_mesh(vparts, rank), cfg, nregs,
                      comm=comm)


def _kinds(sysm):
    return [getattr(k, 'kind', None) for g in sysm.rhs_graphs(0, 1)
            for w, k in g.plan if w == 'kernel']


def _slow(*case):
    """An opt-in kernel variant that is not a default (PYFR_B200_SLOW=1)"""
    return pytest.param(*case, marks=pytest.mark.slow)


@pytest.mark.parametrize('kw,opts,expect', [
    (dict(order=2, warp=0.1), {}, 'gradflux'),
    (dict(order=2), {}, 'gradflux'),                           # affine path
    (dict(order=3, rsolver='hllc', beta=0.0, warp=0.1), {}, 'gradflux'),
    (dict(order=2, beta=-0.5, curved=0.5, warp=0.1), {}, 'gradflux'),
    (dict(order=2, warp=0.1), {'fusion': 0}, 'tflux'),
    (dict(order=2, warp=0.1), {'dead-rows': 0, 'gradflux-monojac': 0},
     'gradflux'),
    _slow(dict(order=2, warp=0.1), {'gradflux-planes': 1}, 'gradflux'),
    (dict(order=3), {'n-soa': 4}, 'gradflux'),
    # two adjacent columns per work item (16-byte accesses)
    _slow(dict(order=2, warp=0.1), {'gradflux-vec2': 'p1,p3,p5'}, 'gradflux'),
    _slow(dict(order=3, beta=0.0), {'gradflux-vec2': 'p3'}, 'gradflux'),
    _slow(dict(order=2, warp=0.1), {'gradflux-vec2': 'p1,p3,p5',
                                    'gradflux-planes': 1}, 'gradflux'),
    _slow(dict(order=4), {'gradflux-vec2': 'p1,p3,p5'}, 'gradflux'),
    # intconu over pairs of points (128-bit accesses where both addresses
    # of a side are adjacent): one-sided, central and left-biased LDG
    _slow(dict(order=2, warp=0.1), {'conu-pairs': 1, 'conu-fold': 0},
          'intconu'),
    _slow(dict(order=3, rsolver='hllc', beta=0.0, warp=0.1),
          {'conu-pairs': 1}, 'intconu'),
    _slow(dict(order=2, beta=-0.5, curved=0.5, warp=0.1),
          {'conu-pairs': 1, 'conu-fold': 0}, 'intconu'),
    _slow(dict(order=2, beta=0.25), {'conu-pairs': 1, 'fusion': 0},
          'intconu'),
    _slow(dict(order=3), {'conu-pairs': 1, 'n-soa': 4, 'conu-fold': 0},
          'intconu'),
    # ... with the interface points in true address order (most pairs
    # then take the 128-bit path)
    _slow(dict(order=2, warp=0.1), {'conu-pairs': 1, 'conu-fold': 0,
                                    'inters-order': 'address'}, 'intconu'),
    _slow(dict(order=3, rsolver='hllc', beta=0.0),
          {'conu-pairs': 1, 'inters-order': 'address'}, 'intconu'),
    _slow(dict(order=2, beta=-0.5, warp=0.1),
          {'conu-pairs': 1, 'conu-fold': 0, 'inters-order': 'address'},
          'intconu'),
    (dict(order=4), {'inters-order': 'address'}, 'gradflux'),
    # the table-driven fused kernel (what non-tensor-product elements and
    # hexes with gradflux-tensor = 0 take)
    (dict(order=2, warp=0.1), {'gradflux-tensor': 0}, 'gradflux'),
    (dict(order=2), {'gradflux-tensor': 0}, 'gradflux'),
    (dict(order=3, rsolver='hllc', beta=0.0, curved=0.5, warp=0.1),
     {'gradflux-tensor': 0}, 'gradflux'),
    (dict(order=4), {'gradflux-tensor': 0}, 'gradflux'),
    # sum-factorised kernel: curved region + linear region, no dead rows
    (dict(order=3, rsolver='hllc', beta=0.0, curved=0.5, warp=0.1), {},
     'gradflux'),
    (dict(order=2, warp=0.1), {'dead-rows': 0}, 'gradflux'),
    (dict(order=1), {}, 'gradflux'),
    # the benchmark's kernel (p = 4): common solution gathered by the
    # element kernel (whole rows by bulk copy, the rest point by point);
    # the fold and the row copies switched off; the opt-in half-block form
    (dict(order=4), {}, 'gradflux'),
    (dict(order=4), {'conu-fold': 0}, 'intconu'),
    (dict(order=4), {'gather-rows': 0}, 'gradflux'),
    (dict(order=4), {'gradflux-split': 1}, 'gradflux'),
    (dict(order=4, beta=-0.5, warp=0.1), {'conu-fold': 0,
                                          'gradflux-split': 1}, 'intconu'),
    (dict(order=4, beta=0.0, warp=0.1), {}, 'intconu'),
    (dict(order=4, beta=-0.5, curved=0.5, warp=0.1), {}, 'gradflux'),
    (dict(order=4, warp=0.1, rsolver='hllc'), {}, 'gradflux'),
], ids=str)
def test_navier_stokes_rhs_through_generated_kernels(emulated, kw, opts,
                                                     expect):
    n = (3, 2, 2)
    cfg, box = cases.make('tgv', n, **kw)
    sysm = _b200(cfg, box, opts=opts)
    sysm.rhs(0.0, 0, 1)
    out = sysm.ele_scal_upts(1)[0]

    _, ref = oracle_rhs('tgv', n, **kw)
    esys, ext = oracle_rhs('tgv', n, extended=True, **kw)

    assert expect in _kinds(sysm)
    assert_parity(out, ref[0], ext[0], 1e-12, mag=rhs_magnitude(esys[0])[0])

    # which fused element kernel ran: the sum-factorised one for hexes
    # unless switched off or the table-driven kernel's own options are set
    for g in sysm.rhs_graphs(0, 1):
        for w, k in g.plan:
            if w == 'kernel' and getattr(k, 'kind', None) == 'gradflux':
                old = any(o.startswith('gradflux-') and o not in
                          ('gradflux-threads', 'gradflux-split')
                          for o in opts)
                assert k.info['tensor'] == (not old), opts


def _vec2_cases():
    from test_gpu_zlate import VEC2_CASES
    return VEC2_CASES


@pytest.mark.slow
@pytest.mark.parametrize('case,n,kw,opts', _vec2_cases(), ids=str)
def test_vectorised_gradflux_device_cases(emulated, case, n, kw, opts):
    """The device parity cases of the gradflux-vec2 variants
    (tests/test_gpu_zlate.py), here on the execution model; the model also
    checks that every 16-byte access is 16-byte aligned."""
    cfg, box = cases.make(case, n, **kw)
    sysm = _b200(cfg, box, opts=opts)
    assert 'gradflux' in _kinds(sysm)
    sysm.rhs(0.0, 0, 1)
    out = sysm.ele_scal_upts(1)[0]

    _, ref = oracle_rhs(case, n, **kw)
    if kw.get('precision') == 'single':
        _, r64 = oracle_rhs(case, n, **{**kw, 'precision': 'double'})
        floor = rel_err(ref[0].astype(float), r64[0])
        assert rel_err(out.astype(float), r64[0]) <= max(4*floor, 1e-5)
    else:
        _, ext = oracle_rhs(case, n, extended=True, **kw)
        assert_parity(out, ref[0], ext[0], 1e-12)


@pytest.mark.parametrize('opts,order', [
    ({}, 2), ({'gradflux-vec2': 'p1,p3,p5', 'conu-pairs': 1,
               'inters-order': 'address'}, 2),
    # BASELINE configs[4] proxy: p = 6, SoA width 4 chosen automatically,
    # the sum-factorised fused kernel (four columns per access)
    ({}, 6)
], ids=['default', 'vec2+pairs', 'p6'])
def test_fp32_kernels(emulated, opts, order):
    n, kw = (3, 2, 2) if order == 2 else (2, 2, 2), dict(order=order,
                                                         warp=0.1)
    cfg, box = cases.make('tgv', n, precision='single', **kw)
    sysm = _b200(cfg, box, opts=opts)
    sysm.rhs(0.0, 0, 1)
    out = sysm.ele_scal_upts(1)[0]

    _, r64 = oracle_rhs('tgv', n, **kw)
    _, r32 = oracle_rhs('tgv', n, precision='single', **kw)
    floor = rel_err(r32[0].astype(float), r64[0])

    assert out.dtype == np.float32
    assert rel_err(out.astype(float), r64[0]) <= max(4*floor, 1e-5)
    if order == 6:
        assert sysm.backend.soasz == 4 and 'gradflux' in _kinds(sysm)


@pytest.mark.parametrize('kw,opts,expect', [
    (dict(order=3), {}, 'fluxdiv'),
    (dict(order=2, rsolver='hllc'), {'euler-fusion': 0}, 'tflux'),
], ids=str)
def test_euler_rhs_through_generated_kernels(emulated, kw, opts, expect):
    cfg, box = cases.make('vortex', 5, **kw)
    sysm = _b200(cfg, box, opts=opts)
    sysm.rhs(0.0, 0, 1)
    out = sysm.ele_scal_upts(1)[0]

    _, ref = oracle_rhs('vortex', 5, **kw)
    _, ext = oracle_rhs('vortex', 5, extended=True, **kw)

    assert expect in _kinds(sysm)
    assert_parity(out, ref[0], ext[0], 1e-12)


@pytest.mark.parametrize('system,n,bcs,kw', [
    ('navier-stokes', (3, 3, 2), {'ylo': 'no-slp-adia-wall',
                                  'yhi': 'char-riem-inv'},
     dict(order=2, warp=0.1)),
    ('navier-stokes', (2, 2, 3), {'xlo': 'sub-in-frv', 'xhi': 'sub-out-fp',
                                  'zlo': 'slp-adia-wall',
                                  'zhi': 'no-slp-isot-wall'},
     dict(order=2, rsolver='hllc')),
    ('euler', (5, 4), {'xlo': 'char-riem-inv', 'xhi': 'sup-out-fn',
                       'ylo': 'slp-adia-wall', 'yhi': 'sup-in-fa'},
     dict(order=2)),
    ('navier-stokes', (3, 2, 2), {'xlo': 'sub-in-ftpttang',
                                  'xhi': 'sup-out-fn'},
     dict(order=2, warp=0.1)),
], ids=str)
def test_boundary_kernels(emulated, system, n, bcs, kw):
    outs = []
    for which in ('oracle', 'oracle-ext', 'b200'):
        cfg, box, _ = cases.box_case(system, n, bcs, **kw)
        if which == 'b200':
            sysm = _b200(cfg, box)
        else:
            cfg.set('backend-oracle', 'extended-mul', which != 'oracle')
            sysm = get_system(OracleBackend(cfg), box.local_mesh(), cfg, 2)
        sysm.rhs(0.5, 0, 1)
        outs.append(sysm.ele_scal_upts(1)[0])

    assert_parity(outs[2], outs[0], outs[1], 1e-12)


class EmuWorld:
    """In-process stand-in for the NCCL communicator: sends park a copy of
    the device buffer, receives are filled when a graph stage closes."""

    def __init__(self, size):
        self.size, self.box, self.pending = size, {}, []

    def peer(self, rank):
        w = self

        class Comm:
            size = w.size

            def exchange(self, reqs, stream):
                # (an operation on the stream: deferred while a graph is
                # being captured, like the NCCL calls it stands in for)
                def op():
                    for r in reqs:
                        m = r.mat
                        nb = m.nrow*m.ncol*m.itemsize
                        if r.kind == 'send':
                            w.box[rank, r.peer, r.tag] = ct.string_at(m.data,
                                                                      nb)
                        else:
                            w.pending.append(((r.peer, rank, r.tag), m.data,
                                              nb))
                self.rt._do(op)

        c = Comm()
        c.rank = rank
        return c

    def deliver(self):
        for key, ptr, nb in self.pending:
            ct.memmove(ptr, self.box.pop(key), nb)
        self.pending.clear()


@pytest.mark.parametrize('graphs', ['false', 'true'])
@pytest.mark.parametrize('kw', [dict(order=2, warp=0.1),
                                dict(order=2, beta=0.0, rsolver='hllc')],
                         ids=str)
def test_partitioned_run_through_generated_kernels(emulated, kw, graphs):
    """Two partitions, halo exchange through the backend's exchange
    descriptors (pack kernels, mpiconu / mpicflux, receive-only graphs):
    against the partitioned oracle."""
    n, parts = (4, 2, 2), (2, 1, 1)
    _, box = cases.make('tgv', n, **kw)
    vparts = box.brick_partition(parts)
    world = EmuWorld(2)

    systems = []
    for r in range(2):
        cfg, box = cases.make('tgv', n, **kw)
        systems.append(_b200(cfg, box, vparts, r, comm=world.peer(r),
                             opts={'graphs': graphs}))

    for stage in zip(*[s.rhs_graphs(0, 1) for s in systems]):
        for g in stage:
            g.run()
        world.deliver()

    _, ref = oracle_rhs('tgv', n, vparts=vparts, nparts=2, **kw)
    _, ext = oracle_rhs('tgv', n, vparts=vparts, nparts=2, extended=True,
                        **kw)

    for r, s in enumerate(systems):
        assert_parity(s.ele_scal_upts(1)[0], ref[r], ext[r], 1e-12)
        assert 'mpiconu' in _kinds(s) and 'copy' not in _kinds(s)


def test_rk4_and_integrals_through_generated_kernels(emulated):
    from pyfr_b200.host.integrator import (FieldIntegrator, RK4Stepper,
                                           TGV_EXPRS)

    res = []
    for which in ('oracle', 'b200'):
        cfg, box = cases.make('tgv', (3, 2, 2), order=2, warp=0.1)
        sysm = (_b200(cfg, box, nregs=3) if which == 'b200' else
                get_system(OracleBackend(cfg), box.local_mesh(), cfg, 3))
        fi, st = FieldIntegrator(sysm, cfg, TGV_EXPRS), RK4Stepper(sysm)
        h = [fi(0.0, st.idxcurr)]
        st.advance(3, 2e-3)
        h.append(fi(st.tcurr, st.idxcurr))
        res.append((np.array(h), st.soln[0]))

    (ho, so), (hb, sb) = res
    assert np.abs(hb/ho - 1).max() < 1e-12
    assert rel_err(sb, so) < 1e-12


def test_standalone_driver(emulated, capsys, monkeypatch):
    """python -m pyfr_b200: RK4 steps + integrals, end to end."""
    from pyfr_b200.__main__ import main

    main(['tgv', '--n', '2', '--order', '2', '--steps', '2', '--every', '1',
          '--opt', 'graphs=false'])
    lines = [l for l in capsys.readouterr().out.splitlines()
             if l and not l.startswith('#')]

    assert len(lines) == 3
    ke0 = float(lines[0].split()[2])
    assert abs(ke0/(2*np.pi)**3 - 0.125) < 2e-2


def test_bench_script_end_to_end(emulated, monkeypatch, capsys):
    """bench.py's whole main path (timed loop, per-kernel timing, roofline
    bookkeeping, pipelined end-to-end section, JSON line) on the emulated
    runtime; event timings are faked, so only the structure is checked."""
    import json
    import runpy

    monkeypatch.setattr(emu.EmuRuntime, 'elapsed_ms', lambda s, a, b: 1.0)
    monkeypatch.setattr(sys, 'argv', [
        'bench.py', '--n', '2', '--order', '2', '--steps', '3', '--warmup',
        '1', '--no-cpu', '--no-clocks', '--no-graphs', '--timestep'
    ])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runpy.run_path(os.path.join(root, 'bench.py'), run_name='__main__')

    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])

    # the partitioned-oracle check that precedes the timing, and the whole
    # time steps timed after it
    assert line['parity']['nranks'] == 1 and line['parity']['err'] < 1e-11
    assert set(line['time_step']) == {'rk4', 'rk45', 'rk45_fused_update'}
    assert 'stage_kernels_ms' in line['time_step']['rk45_fused_update']

    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup',
                'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
                'dtype', 'data', 'config', 'gpu_launches', 'roofline', 'e2e',
                'clocks', 'cpu_baseline'):
        assert key in line, key

    assert line['launches_per_step'] == 4 and line['gpu_launches'] == 12
    assert line['e2e']['h2d_bytes_per_step'] == 27*5*8*8
    assert set(line['roofline']) >= {'bound', 'achieved', 'peak', 'unit',
                                     'frac', 'traffic'}


def test_bench_script_with_opt_in_variants(emulated, monkeypatch, capsys,
                                           tmp_path):
    """The invocation the device suite's variant timing report makes
    (tests/test_gpu_zlate.py): bench.py with --opt switches and
    --kernel-times."""
    import json
    import runpy

    from test_gpu_zlate import VARIANT_REPORT

    monkeypatch.setattr(emu.EmuRuntime, 'elapsed_ms', lambda s, a, b: 1.0)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    kt = str(tmp_path / 'kt.json')

    tag, opts = VARIANT_REPORT[-1]
    argv = ['bench.py', '--n', '2', '--order', '2', '--steps', '2',
            '--warmup', '1', '--no-cpu', '--no-e2e', '--no-clocks',
            '--no-graphs', '--kernel-times', kt]
    for o in opts:
        argv += ['--opt', o]
    monkeypatch.setattr(sys, 'argv', argv)
    runpy.run_path(os.path.join(root, 'bench.py'), run_name='__main__')

    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line['value'] > 0 and line['launches_per_step'] == 4
    with open(kt) as f:
        kern = json.load(f)['kernels']
    assert any(k.endswith('gradflux') for k in kern)
    assert all('ms' in v for v in kern.values())


@pytest.mark.parametrize('case,n,kw', [
    ('tgv', (3, 2, 2), dict(order=2, warp=0.1, antialias='surf-flux')),
    ('tgv', (3, 2, 2), dict(order=2, antialias='flux, surf-flux',
                            beta=0.0)),
    ('vortex', 5, dict(order=3, antialias='surf-flux', rsolver='hllc')),
], ids=str)
def test_surface_flux_antialiasing(emulated, case, n, kw):
    """Surface-flux anti-aliasing: more flux points than the polynomial
    degree needs, the projection folded into M3 / M6."""
    cfg, box = cases.make(case, n, **kw)
    sysm = _b200(cfg, box)
    sysm.rhs(0.0, 0, 1)
    out = sysm.ele_scal_upts(1)[0]

    _, ref = oracle_rhs(case, n, **kw)
    _, ext = oracle_rhs(case, n, extended=True, **kw)
    _, noaa = oracle_rhs(case, n, **{**kw, 'antialias': 'none'})

    assert_parity(out, ref[0], ext[0], 1e-12)
    assert rel_err(ref[0], noaa[0]) > 1e-5


@pytest.mark.parametrize('case,n,kw', [
    ('tgv', (3, 2, 2), dict(order=2, warp=0.1, antialias='flux')),
    ('vortex', 5, dict(order=3, antialias='flux', rsolver='hllc')),
], ids=str)
def test_flux_antialiasing_through_generated_kernels(emulated, case, n, kw):
    """Flux anti-aliasing (solution interpolated to quadrature points with
    M7, flux evaluated there, divergence through (M1 - M3*M2)*M9): generic
    operator and flux kernels on other point sets."""
    cfg, box = cases.make(case, n, **kw)
    sysm = _b200(cfg, box)
    sysm.rhs(0.0, 0, 1)
    out = sysm.ele_scal_upts(1)[0]

    _, ref = oracle_rhs(case, n, **kw)
    _, ext = oracle_rhs(case, n, extended=True, **kw)
    _, noaa = oracle_rhs(case, n, **{**kw, 'antialias': 'none'})

    assert_parity(out, ref[0], ext[0], 1e-12)
    assert rel_err(ref[0], noaa[0]) > 1e-4          # it does something
    assert 'tflux' in _kinds(sysm) and 'gradflux' not in _kinds(sysm)


def test_sutherland_viscosity_through_generated_kernels(emulated):
    n, kw = (3, 2, 2), dict(order=2, warp=0.1, visc_corr='sutherland')
    cfg, box = cases.make('tgv', n, **kw)
    sysm = _b200(cfg, box)
    sysm.rhs(0.0, 0, 1)
    out = sysm.ele_scal_upts(1)[0]

    _, ref = oracle_rhs('tgv', n, **kw)
    _, ext = oracle_rhs('tgv', n, extended=True, **kw)
    _, const = oracle_rhs('tgv', n, **{**kw, 'visc_corr': 'none'})

    assert_parity(out, ref[0], ext[0], 1e-12)
    assert rel_err(ref[0], const[0]) > 1e-7


def _pi_run(sysm, cfg, tend, norm='l2', fused=False):
    from pyfr_b200.host.integrator import PIController, RK45Stepper

    cfg.set('solver-time-integrator', 'dt', 0.05)
    cfg.set('solver-time-integrator', 'atol', 1e-6)
    cfg.set('solver-time-integrator', 'rtol', 1e-6)
    cfg.set('solver-time-integrator', 'errest-norm', norm)

    st = RK45Stepper(sysm, errest=True, fused=fused)
    pi = PIController(st, cfg, ['rho', 'rhou', 'rhov', 'E'])
    pi.advance_to(tend)

    return pi, st


@pytest.mark.parametrize('norm', ['l2', 'uniform'])
def test_rk45_pi_controller_through_generated_kernels(emulated, norm):
    """BASELINE configs[0]: Euler vortex, rk45 + PI controller.  The
    rkvdh2 and reduction kernels and the accept/reject history must agree
    with the oracle backend driven by the same host code."""
    res = []
    for which in ('oracle', 'b200'):
        cfg, box = cases.make('vortex', (4, 4), order=3)
        sysm = (_b200(cfg, box, nregs=4) if which == 'b200' else
                get_system(OracleBackend(cfg), box.local_mesh(), cfg, 4))
        # the l2 case also takes the fused stage update (dt rebound at
        # every step, rejected steps included)
        pi, st = _pi_run(sysm, cfg, 0.2, norm,
                         fused=which == 'b200' and norm == 'l2')
        res.append((pi.stepinfo, st.soln[0], pi))

    (io, so, po), (ib, sb, pb) = res
    assert [a[1] for a in io] == [a[1] for a in ib]
    assert po.nacptsteps >= 3 and po.nrjctsteps >= 1
    assert pb.tcurr == po.tcurr == 0.2
    np.testing.assert_allclose([a[0] for a in ib], [a[0] for a in io],
                               rtol=1e-9)
    np.testing.assert_allclose([a[2] for a in ib], [a[2] for a in io],
                               rtol=1e-8)
    assert rel_err(sb, so) < 1e-12


def test_reduction_kernel_ignores_padding(emulated):
    """Elements beyond neles in the last block hold whatever the kernels
    left there; the reduction must not see them."""
    cfg, box = cases.make('vortex', (3, 3), order=2)
    sysm = _b200(cfg, box, nregs=2)
    be = sysm.backend
    a, b = sysm.ele_banks[0]

    rng = np.random.default_rng(3)
    va, vb = (rng.standard_normal(a.ioshape) for _ in range(2))
    a.set(va)
    b.set(vb)

    # poison the padding through the raw storage image
    raw = np.empty(a.nbytes // a.itemsize)
    be.rt.memcpy(raw.ctypes.data, a.data, a.nbytes)
    be.rt.device_sync()
    neles = a.ioshape[-1]
    assert neles % be.csubsz, 'case must have a ragged last block'
    marker = raw.copy()
    a.set(np.full(a.ioshape, 7.0))
    be.rt.memcpy(marker.ctypes.data, a.data, a.nbytes)
    be.rt.device_sync()
    raw[marker != 7.0] = 1e30
    be.rt.memcpy(a.data, raw.ctypes.data, a.nbytes)
    be.rt.device_sync()

    pv = (0.5, 1.0, 2.0, 4.0)
    for rop, red in (('sum', np.sum), ('max', np.max)):
        k = be.kernel('reduction', rop, ['s*x*y + w', 'fabs(x)'],
                      {'x': a, 'y': b}, svars=['s'], pvars={'w': pv})
        be.commit()
        k.bind(1.5)
        be.run_kernels([k], wait=True)

        w = np.array(pv)[None, :, None]
        want = [red(1.5*va*vb + w), red(np.abs(va))]
        np.testing.assert_allclose(k.retval, want, rtol=1e-13)


def test_standalone_driver_adaptive(emulated, capsys):
    """python -m pyfr_b200 vortex --scheme rk45: PI-controlled run."""
    from pyfr_b200.__main__ import main

    main(['vortex', '--n', '3', '--order', '2', '--scheme', 'rk45', '--dt',
          '0.05', '--steps', '4', '--every', '2', '--fused-update'])
    out = capsys.readouterr().out.splitlines()
    rows = [l.split() for l in out if l and not l.startswith('#')]

    assert len(rows) == 3 and float(rows[-1][1]) == pytest.approx(0.2)
    # mass is conserved by the scheme to round-off
    assert abs(float(rows[-1][2])/float(rows[0][2]) - 1) < 1e-13
    assert 'accepted' in out[-1]


@pytest.mark.parametrize('case,n,kw', [
    ('vortex', (4, 3), dict(order=3)),
    ('tgv', (2, 2, 3), dict(order=2, warp=0.1)),
    ('tgv', (3, 2, 2), dict(order=2, warp=0.1, curved=0.5)),
], ids=['linear', 'curved', 'mixed'])
def test_wavespeed_and_cfl_controller(emulated, case, n, kw):
    """wavespeed kernel, max reduction and the CFL step-size rule against
    the oracle backend under the same host code."""
    from pyfr_b200.host.integrator import CFLController, RK45Stepper

    res, tend = [], None
    for which in ('oracle', 'b200'):
        cfg, box = cases.make(case, n, **kw)
        if which == 'b200':
            cfg.set('backend-b200', 'graphs', 'false')
            from pyfr_b200.backend import B200Backend
            be = B200Backend(cfg)
        else:
            be = OracleBackend(cfg)
        sysm = get_system(be, box.local_mesh(), cfg, 2, needs_cfl=True)

        for k, v in (('dt', 0.01), ('cfl', 0.5), ('cfl-nsteps', 2)):
            cfg.set('solver-time-integrator', k, v)
        lam = sysm.compute_max_wavespeed(0)
        ctl = CFLController(RK45Stepper(sysm), cfg)
        tend = tend or 3.5*ctl._compute_dt_cfl(0)
        ctl.advance_to(tend)
        res.append((lam, [d for d, *_ in ctl.stepinfo],
                    ctl.stepper.soln[0]))

    (lo, do, so), (lb, db, sb) = res
    assert lb == pytest.approx(lo, rel=1e-13)
    assert len(do) >= 3
    np.testing.assert_allclose(db, do, rtol=1e-12)
    assert rel_err(sb, so) < 1e-12


def test_standalone_driver_cfl(emulated, capsys):
    from pyfr_b200.__main__ import main

    main(['tgv', '--n', '2', '--order', '2', '--cfl', '0.4', '--dt', '0.01',
          '--steps', '4', '--every', '4', '--opt', 'graphs=false'])
    out = capsys.readouterr().out.splitlines()
    rows = [l.split() for l in out if l and not l.startswith('#')]

    assert float(rows[-1][1]) == pytest.approx(0.04)
    assert int(rows[-1][0]) >= 2 and 'accepted' in out[-1]


def _mixed_outs(pattern, n, kw, b200):
    outs = []
    for which in ('oracle', 'oracle-ext', 'b200'):
        cfg, box, _ = cases.mixed_case(pattern, n, **kw)
        if which == 'b200':
            sysm = b200(cfg, box)
        else:
            cfg.set('backend-oracle', 'extended-mul', which != 'oracle')
            sysm = get_system(OracleBackend(cfg), box.local_mesh(), cfg, 2)
        sysm.rhs(0.0, 0, 1)
        if which == 'b200':
            sysm.backend.wait()
        outs.append(sysm.ele_scal_upts(1))

    return outs, sysm


@pytest.mark.parametrize('pattern,n,kw,nfused', [
    ('quad+tri', (4, 3), dict(order=3, rsolver='hllc'), 2),
    ('hex+pri', (3, 2, 2), dict(order=2, beta=0.0), 2),
    ('hex+pri+pyr+tet', (4, 2, 2), dict(order=3), 2),
], ids=str)
def test_mixed_element_types(emulated, pattern, n, kw, nfused):
    """BASELINE configs[3] through the host mirror: several element types
    in one mesh, mixed-face interface views, dense operators (tabulated
    shapes, pyfr_b200/host/data)."""
    (ref, ext, out), sysm = _mixed_outs(pattern, n, kw, _b200)

    assert len(out) == len(pattern.split('+'))
    for o, r, e in zip(out, ref, ext):
        assert_parity(o, r, e, 1e-12)

    kinds = _kinds(sysm)
    assert kinds.count('fluxdiv') + kinds.count('gradflux') == nfused


@pytest.mark.slow
def test_dense_operators_from_constant_table(emulated):
    """mul-const-table: operators with many distinct coefficients (tets,
    pyramids) read them from __constant__ memory instead of literals."""
    import functools

    b200 = functools.partial(_b200, opts={'mul-const-table': 32})
    (ref, ext, out), sysm = _mixed_outs('hex+pri+pyr+tet', (4, 2, 2),
                                        dict(order=3), b200)
    for o, r, e in zip(out, ref, ext):
        assert_parity(o, r, e, 1e-12)

    muls = [k for g in sysm.rhs_graphs(0, 1) for w, k in g.plan
            if w == 'kernel' and (k.kind or '').startswith('mul')]
    assert muls


def test_bench_script_mixed_case(emulated, monkeypatch, capsys):
    """bench.py --case hex+pri: the configs[3]-style diagnostic mode."""
    import json
    import runpy

    monkeypatch.setattr(emu.EmuRuntime, 'elapsed_ms', lambda s, a, b: 1.0)
    monkeypatch.setattr(sys, 'argv', [
        'bench.py', '--case', 'hex+pri', '--n', '2', '--order', '2',
        '--steps', '2', '--warmup', '1', '--no-clocks', '--no-graphs'
    ])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runpy.run_path(os.path.join(root, 'bench.py'), run_name='__main__')

    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert 'hex+pri' in line['config']['workload']
    assert line['value'] > 0 and line['rhs_model'] is None
    # 4 hexes (27 points) and 8 prisms (18 points), 5 variables
    assert line['dof'] == (4*27 + 8*18)*5


@pytest.mark.parametrize('case,n,kw', [
    ('tgv', (3, 2, 2), dict(order=2, warp=0.1)),
    ('tgv', (3, 2, 2), dict(order=2, warp=0.1, beta=0.0, rsolver='hllc')),
    ('vortex', 5, dict(order=3)),
], ids=str)
def test_int64_index_model(emulated, case, n, kw):
    """``[backend] memory-model = large``: 64-bit view / packing indices
    (pyfr/backends/base/backend.py:60-68)."""
    cfg, box = cases.make(case, n, **kw)
    cfg.set('backend', 'memory-model', 'large')
    sysm = _b200(cfg, box)
    assert sysm.backend.ixdtype == np.int64
    sysm.rhs(0.0, 0, 1)
    out = sysm.ele_scal_upts(1)[0]

    _, ref = oracle_rhs(case, n, **kw)
    _, ext = oracle_rhs(case, n, extended=True, **kw)
    assert_parity(out, ref[0], ext[0], 1e-12)


def test_field_reductions_with_coordinates(emulated):
    """fieldeval beyond the weighted sum: coordinate-dependent integrands,
    per-element maxima and minima (L-inf norms of the integrate plugin)."""
    from pyfr_b200.host.integrator import FieldIntegrator

    exprs = ['rho*x*x + p*cos(y)', 'fabs(u - 0.1*z) + grad_v_x*t']
    res = {}
    for which in ('oracle', 'b200'):
        for rop in ('sum', 'max', 'min'):
            cfg, box = cases.make('tgv', (3, 2, 2), order=2, warp=0.1)
            sysm = (_b200(cfg, box) if which == 'b200' else
                    get_system(OracleBackend(cfg), box.local_mesh(), cfg, 2))
            fi = FieldIntegrator(sysm, cfg, exprs, reduceop=rop)
            res[which, rop] = fi(0.3, 0)

    for rop in ('sum', 'max', 'min'):
        np.testing.assert_allclose(res['b200', rop], res['oracle', rop],
                                   rtol=1e-12)

    # independent check of the maxima from the solution itself
    cfg, box = cases.make('tgv', (3, 2, 2), order=2, warp=0.1)
    sysm = get_system(OracleBackend(cfg), box.local_mesh(), cfg, 2)
    u, x = sysm.ele_scal_upts(0)[0], sysm.ele_ploc_upts[0]()
    rho, E = u[:, 0], u[:, 4]
    p = 0.4*(E - 0.5*(u[:, 1]**2 + u[:, 2]**2 + u[:, 3]**2)/rho)
    want = (rho*x[:, 0]**2 + p*np.cos(x[:, 1])).max()
    assert res['oracle', 'max'][0] == pytest.approx(want, rel=1e-13)
    assert res['oracle', 'min'][0] < res['oracle', 'max'][0]


@pytest.mark.parametrize('case,n,kw,kind,errest', [
    ('tgv', (3, 2, 2), dict(order=2, warp=0.1), 'mul+negdivconf+rkvdh2',
     False),
    ('vortex', (4, 4), dict(order=3), 'fluxdiv+rkvdh2', False),
    ('vortex', (4, 4), dict(order=3), 'fluxdiv+rkvdh2', True),
    ('tgv', (3, 2, 2), dict(order=2, warp=0.1, curved=0.5),
     'mul+negdivconf+rkvdh2', True),
], ids=str)
def test_rk_stage_update_fused_into_last_rhs_kernel(emulated, case, n, kw,
                                                    kind, errest):
    """the rkvdh2 stage update applied in the epilogue of
    the last RHS kernel.  Same result as the separate kernels and as the
    oracle; one launch fewer per stage."""
    from pyfr_b200.host.integrator import RK45Stepper

    sols, launches = {}, {}
    for which in ('oracle', 'b200', 'b200-fused'):
        cfg, box = cases.make(case, n, **kw)
        sysm = (get_system(OracleBackend(cfg), box.local_mesh(), cfg, 4)
                if which == 'oracle' else _b200(cfg, box, nregs=4))
        st = RK45Stepper(sysm, errest=errest, fused=which != 'b200')
        st.advance(1, 2e-3)                      # builds the kernels
        rt = getattr(sysm.backend, 'rt', None)
        n0 = getattr(rt, 'nlaunch', 0)
        st.advance(2, 2e-3)
        launches[which] = getattr(rt, 'nlaunch', 0) - n0
        sols[which] = [sysm.ele_scal_upts(i)[0] for i in range(4)]

        if which == 'b200-fused':
            kinds = [getattr(k, 'kind', None) for g in sysm._graphs.values()
                     for gg in g for w, k in gg.plan if w == 'kernel']
            assert kind in kinds and 'rkvdh2' not in kinds

    # every register bank that carries state (solution, previous solution,
    # error estimate) agrees; the bank left holding a raw RHS is not state
    # (the error estimate is a cancelling sum of O(dt k) terms: measure it
    # against the scale of the solution it estimates the error of)
    live = [st.idxcurr] + (sorted(set(range(4)) - {0, 1}) if errest else [])
    scale = np.abs(sols['oracle'][st.idxcurr]).max()
    for i in live:
        for other, tol in (('oracle', 1e-12), ('b200', 1e-13)):
            d = np.abs(sols['b200-fused'][i] - sols[other][i]).max()
            assert d < tol*scale, (i, other, d)

    # 5 stages x 2 steps, one launch saved per stage and element region
    assert launches['b200'] - launches['b200-fused'] >= 10


def _b200_graphs(cfg, box, nregs=2, **kw):
    from pyfr_b200.backend import B200Backend

    cfg.set('backend-b200', 'graphs', 'true')
    be = B200Backend(cfg)
    assert be.use_graphs and getattr(be.rt, 'emulated', False)
    return get_system(be, box.local_mesh(), cfg, nregs, **kw)


def test_graph_replay_without_recapture(emulated):
    """``graphs = true``: each RHS graph is captured once and replayed.  The
    emulated runtime copies kernel parameters by value at capture time, as
    CUDA graphs do, so a run-time scalar that changes (here the time ``t``
    a boundary condition depends on) is only seen because the kernels read
    it from the backend's device-resident scalar block, which is refreshed
    before every launch: the graphs are never captured again."""
    bcs = {'xlo': 'sub-in-frv', 'xhi': 'sub-out-fp'}
    tdep = '\n[soln-bcs-xlo]\nu = 0.2 + 0.1*sin(3*t)\n'

    def systems():
        out = []
        for which in ('oracle', 'b200'):
            cfg, box, txt = cases.box_case('navier-stokes', (3, 2, 2), bcs,
                                           order=2, warp=0.1)
            cfg.set('soln-bcs-xlo', 'u', '0.2 + 0.1*sin(3*t)')
            out.append(_b200_graphs(cfg, box) if which == 'b200' else
                       get_system(OracleBackend(cfg), box.local_mesh(), cfg,
                                  2))
        return out

    so, sb = systems()
    rt = sb.backend.rt

    outs = []
    for t in (0.0, 0.0, 0.4, 0.4, 0.9):
        for s in (so, sb):
            s.rhs(t, 0, 1)
        sb.backend.wait()
        outs.append((t, so.ele_scal_upts(1)[0], sb.ele_scal_upts(1)[0],
                     rt.ncaptures))

    for t, ro, rb, _ in outs:
        assert rel_err(rb, ro) < 1e-12, t

    # the boundary value really depends on t ...
    assert rel_err(outs[2][1], outs[0][1]) > 1e-6
    # ... and no graph was captured a second time
    ncap = [c for *_, c in outs]
    ngraphs = len(sb.rhs_graphs(0, 1))
    assert ncap == [ngraphs]*len(outs)


def test_fused_stage_update_under_graphs(emulated):
    """PI-controlled RK45 with the fused stage update and graphs on: dt
    changes every step and reaches the captured kernels through the
    device-resident scalar block (no re-capture)."""
    res, rts = [], []
    for which in ('oracle', 'b200'):
        cfg, box = cases.make('vortex', (4, 4), order=3)
        sysm = (_b200_graphs(cfg, box, nregs=4) if which == 'b200' else
                get_system(OracleBackend(cfg), box.local_mesh(), cfg, 4))
        pi, st = _pi_run(sysm, cfg, 0.2, fused=which == 'b200')
        res.append((pi.stepinfo, st.soln[0]))
        rts.append(getattr(sysm.backend, 'rt', None))
        ngraphs = sum(len(gs) for gs in sysm._graphs.values())

    # every graph was captured exactly once, however often dt changed
    assert len(res[1][0]) > 3 and rts[1].ncaptures == ngraphs

    (io, so), (ib, sb) = res
    assert [a[1] for a in io] == [a[1] for a in ib]
    np.testing.assert_allclose([a[0] for a in ib], [a[0] for a in io],
                               rtol=1e-9)
    assert rel_err(sb, so) < 1e-12


@pytest.mark.parametrize('case,n,kw,opts', [
    ('tgv', (3, 2, 2), dict(order=2, warp=0.1), {}),
    ('tgv', (3, 2, 2), dict(order=2, warp=0.1), {'fusion': 0}),
    ('tgv', (3, 2, 2), dict(order=3, warp=0.1, beta=0.0, rsolver='hllc'),
     {}),
    ('tgv', (3, 2, 2), dict(order=2), {}),
    ('vortex', 5, dict(order=3), {}),
    ('vortex', 5, dict(order=3), {'fusion': 0}),
], ids=str)
def test_gauss_lobatto_points(emulated, case, n, kw, opts):
    """Flux points that coincide with solution points (SURVEY appendix B,
    hex GLL row): ``M0`` is a selection, the common solution lives in its
    own buffer and interface views address solution-point rows."""
    kw = dict(kw, pts='gauss-legendre-lobatto')
    cfg, box = cases.make(case, n, **kw)
    sysm = _b200(cfg, box, opts=opts)
    sysm.rhs(0.0, 0, 1)
    out = sysm.ele_scal_upts(1)[0]

    _, ref = oracle_rhs(case, n, **kw)
    _, ext = oracle_rhs(case, n, extended=True, **kw)
    assert_parity(out, ref[0], ext[0], 1e-12)


def test_standalone_driver_rk45_cfl_fused(emulated, capsys):
    """The combination the prepared GPU job times (tools/gpu/r02a.sh)."""
    from pyfr_b200.__main__ import main

    res = []
    for extra in ([], ['--fused-update']):
        main(['tgv', '--n', '2', '--order', '2', '--scheme', 'rk45', '--cfl',
              '0.3', '--dt', '0.01', '--steps', '2', '--every', '2', *extra])
        out = capsys.readouterr().out.splitlines()
        res.append([l.split() for l in out if l and not l.startswith('#')])

    assert res[0][-1][:2] == res[1][-1][:2]            # same steps, same time
    np.testing.assert_allclose([float(v) for v in res[1][-1][2:]],
                               [float(v) for v in res[0][-1][2:]],
                               rtol=1e-12)


@pytest.mark.parametrize('pattern,n,parts,kw', [
    ('quad+tri', (4, 4), (2, 2), dict(order=3)),
    ('hex+pri+pyr+tet', (4, 2, 2), (2, 1, 1), dict(order=2, beta=0.0)),
], ids=str)
def test_partitioned_mixed_mesh(emulated, pattern, n, parts, kw):
    """Mixed element types across partitions: halo views over several
    element types, on the oracle (against the unpartitioned run: with a
    central LDG flux or none at all the RHS does not depend on the
    partitioning) and through the generated kernels."""
    from oracle.npbackend import LocalComm
    from util import run_lockstep

    nparts = int(np.prod(parts))
    cfg, box, _ = cases.mixed_case(pattern, n, **kw)
    vparts = box.brick_partition(parts)
    order = box.partition_order(vparts)

    # unpartitioned oracle
    whole = get_system(OracleBackend(cfg), box.local_mesh(), cfg, 2)
    whole.rhs(0.0, 0, 1)
    wrhs = dict(zip(box.etypes, whole.ele_scal_upts(1)))

    # partitioned oracle
    lworld = LocalComm(0, nparts)
    osys = []
    for r in range(nparts):
        cfg, box, _ = cases.mixed_case(pattern, n, **kw)
        osys.append(get_system(OracleBackend(cfg), box.local_mesh(vparts, r),
                               cfg, 2, comm=lworld.peer(r)))
    run_lockstep(osys, lworld, 0.0, 0, 1)

    # partitioned, generated kernels
    eworld = EmuWorld(nparts)
    bsys = []
    for r in range(nparts):
        cfg, box, _ = cases.mixed_case(pattern, n, **kw)
        bsys.append(_b200(cfg, box, vparts, r, comm=eworld.peer(r)))
    run_lockstep(bsys, eworld, 0.0, 0, 1)

    scale = max(np.abs(a).max() for a in wrhs.values())
    for r in range(nparts):
        ets = [et for et in box.etypes if et in order[r]]
        for et, o, b in zip(ets, osys[r].ele_scal_upts(1),
                            bsys[r].ele_scal_upts(1)):
            assert np.abs(o - wrhs[et][..., order[r][et]]).max() < 1e-12*scale
            assert np.abs(b - o).max() < 1e-12*scale


def test_eight_partitions_as_in_the_scaling_run(emulated):
    """The 2 x 2 x 2 brick layout bench.py uses at 8 GPUs: every rank has
    three neighbours and meets each of them across *two* opposite faces
    (periodic box, two bricks per axis).  Generated kernels with graphs on,
    against the partitioned oracle."""
    n, parts, kw = (4, 4, 4), (2, 2, 2), dict(order=1, warp=0.05)
    _, box = cases.make('tgv', n, **kw)
    vparts = box.brick_partition(parts)
    world = EmuWorld(8)

    systems = []
    for r in range(8):
        cfg, box = cases.make('tgv', n, **kw)
        systems.append(_b200(cfg, box, vparts, r, comm=world.peer(r),
                             opts={'graphs': 'true'}))
        assert sorted(box.local_mesh(vparts, r).con_p) == sorted(
            {r ^ 1, r ^ 2, r ^ 4})

    for _ in range(2):                           # capture, then replay
        for stage in zip(*[s.rhs_graphs(0, 1) for s in systems]):
            for g in stage:
                g.run()
            world.deliver()

    _, ref = oracle_rhs('tgv', n, vparts=vparts, nparts=8, **kw)
    _, ext = oracle_rhs('tgv', n, vparts=vparts, nparts=8, extended=True,
                        **kw)
    for r, s in enumerate(systems):
        assert_parity(s.ele_scal_upts(1)[0], ref[r], ext[r], 1e-12)


def test_operator_kernel_on_random_matrices(emulated):
    """``opmul`` for operators the solvers never hand over: random shapes
    and sparsity (empty rows and columns, fully dense), alpha / beta
    variants, ragged last block, and a shared-memory budget small enough to
    force the input tile to be streamed in several chunks."""
    from pyfr_b200.backend import B200Backend
    from pyfr_b200.host.config import Config

    rng = np.random.default_rng(11)
    cfg = Config('[backend]\nprecision = double\n[backend-b200]\n'
                 'graphs = false\n')
    be = B200Backend(cfg)
    be.smem_budget = 12*1024            # K*LD*8 bytes exceeds this early

    for trial in range(14):
        M, K = int(rng.integers(1, 40)), int(rng.integers(1, 70))
        dens = [0.05, 0.3, 1.0][trial % 3]
        A = rng.standard_normal((M, K))*(rng.random((M, K)) < dens)
        if trial % 4 == 0 and M > 2 and K > 2:
            A[rng.integers(M)] = 0
            A[:, rng.integers(K)] = 0
        alpha = [1.0, -2.0, 0.5][trial % 3]
        beta = [0.0, 1.0, -0.5, 0.0][trial % 4]
        nv, ne = int(rng.integers(1, 6)), int(rng.integers(1, 30))

        b = rng.standard_normal((K, nv, ne))
        c = rng.standard_normal((M, nv, ne))
        ma = be.const_matrix(A)
        mb, mc = be.matrix(b.shape, b, tags={'align'}), \
            be.matrix(c.shape, c, tags={'align'})
        be.commit()

        k = be.kernel('mul', ma, mb, out=mc, alpha=alpha, beta=beta)
        be.run_kernels([k], wait=True)

        want = alpha*np.einsum('mk,kve->mve', A, b) + beta*c
        assert np.abs(mc.get() - want).max() <= 1e-13*max(
            1.0, np.abs(want).max()), (trial, M, K, dens, alpha, beta)


def test_execution_model_traps_misaligned_vector_access():
    """Negative control for the alignment checking of the execution model:
    a 16-byte load at an address that is only 8-byte aligned (a device
    fault) must end the process; the aligned one must not."""
    import subprocess

    prog = r'''
import ctypes as ct, sys
sys.path.insert(0, sys.argv[1])
import emu
src = """
typedef double fpdtype_t;
typedef double2 fpdtype2_t;
extern "C" __global__ void probe(const fpdtype_t* __restrict__ a, fpdtype_t* __restrict__ b, int off)
{
    const fpdtype2_t v = *reinterpret_cast<const fpdtype2_t *>(a + off);
    b[0] = v.x + v.y;
}
"""
rt = emu.EmuRuntime()
a, b = rt.malloc(64), rt.malloc(64)
m = rt.module_load(src.encode())
off = ct.c_int(int(sys.argv[2]))
pa, pb = ct.c_void_p(a), ct.c_void_p(b)
argv = (ct.c_void_p*3)(ct.addressof(pa), ct.addressof(pb), ct.addressof(off))
rt.launch(m, 1, 1, 1, 1, 1, 1, 0, 0, argv)
print('RAN')
'''
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'cudaemu')
    run = lambda off: subprocess.run([sys.executable, '-c', prog, here,
                                      str(off)], capture_output=True,
                                     text=True, timeout=300)
    ok, bad = run(2), run(1)
    assert ok.returncode == 0 and 'RAN' in ok.stdout, ok.stderr[-2000:]
    assert bad.returncode != 0 and 'RAN' not in bad.stdout
