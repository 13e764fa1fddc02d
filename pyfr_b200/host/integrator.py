"""Explicit time stepping and the ``integrate`` post-step reduction.

Host-side counterparts of the callers either side of the RHS path
(SURVEY.md section 8f, ranks 1 and 2):

* ``RK4Stepper`` -- ``pyfr/integrators/explicit/steppers.py:60-108``: four
  RHS evaluations and six ``axnpby`` register updates per step over three
  register banks, with the reference's bank rotation;
* ``FieldIntegrator`` -- ``pyfr/plugins/fieldeval.py:25-248``
  (``BackendFieldReducer`` with ``reduceop='sum'`` at the solution points)
  as used by the ``integrate`` plugin (``pyfr/plugins/integrate.py``):
  expressions over primitive variables, their gradients and coordinates
  are compiled to C (``compile_expr``, ``fieldeval.py:12-22``), evaluated
  per point on the device, weighted with ``w_p |J|`` and summed per
  element by the backend's ``fieldeval`` kernel; the host adds up elements.

* ``RK45Stepper`` + ``PIController`` -- the integrator of BASELINE.json's
  configs[0] (``scheme = rk45``, ``controller = pi``):
  ``pyfr/integrators/explicit/steppers.py:111-243`` (two-register van der
  Houwen scheme; one ``rkvdh2`` kernel per stage, two extra registers for
  the previous solution and the embedded error estimate) and
  ``pyfr/integrators/controllers.py:45-121`` /
  ``pyfr/integrators/explicit/controllers.py:66-131`` (weighted error norm
  by the backend's ``reduction`` kernel, PI step-size law, accept/reject).

Everything numerical is a backend kernel; this file only sequences them.
"""

import math
import re

import numpy as np


class RK4Stepper:
    nregs = 3

    def __init__(self, system, tstart=0.0):
        if system.nrhs < self.nregs:
            raise ValueError('RK4 needs three register banks')

        self.system, self.backend = system, system.backend
        self.tcurr, self.nsteps = tstart, 0
        self.idxcurr = 0
        self._regidx = [0, 1, 2]
        self._addk = {}

    def _add(self, *args):
        consts, regs = args[::2], args[1::2]

        if regs not in self._addk:
            self._addk[regs] = [
                self.backend.kernel('axnpby', *[eb[r] for r in regs])
                for eb in self.system.ele_banks
            ]

        for k in self._addk[regs]:
            k.bind(*consts)

        self.backend.run_kernels(self._addk[regs])

    def _step(self, t, dt):
        add, rhs = self._add, self.system.rhs
        r0, r1, r2 = self._regidx

        if r0 != self.idxcurr:
            r0, r1 = r1, r0

        rhs(t, r0, r1)

        add(0.0, r2, 1.0, r0, dt/2.0, r1)
        rhs(t + dt/2.0, r2, r2)

        add(dt/6.0, r1, 1.0, r0, dt/3.0, r2)

        add(dt/2.0, r2, 1.0, r0)
        rhs(t + dt/2.0, r2, r2)

        add(1.0, r1, dt/3.0, r2)

        add(dt, r2, 1.0, r0)
        rhs(t + dt, r2, r2)

        add(1.0, r1, dt/6.0, r2)

        return r1

    def step(self, dt):
        self.idxcurr = self._step(self.tcurr, dt)
        self.tcurr += dt
        self.nsteps += 1

        return self.idxcurr

    def advance(self, nsteps, dt):
        for _ in range(nsteps):
            self.step(dt)

    @property
    def soln(self):
        return self.system.ele_scal_upts(self.idxcurr)


class RKVdH2RStepper:
    """Low-storage RK schemes of Kennedy, Carpenter & Lewis (2000) in van
    der Houwen form: stage ``i`` evaluates the RHS into ``r2`` and one
    pointwise kernel forms ``r1 + dt a_i r2`` and ``r1 + dt b_i r2``."""

    a, b, bhat = [], [], []
    stepper_order = None

    def __init__(self, system, tstart=0.0, errest=False, fused=False):
        """``fused``: hand each stage's update kernel to the system as a
        post-RHS kernel, grouped with the last RHS kernel, so that the
        backend may apply it in that kernel's epilogue (SURVEY.md section
        8f, rank 1) -- the RHS value then never makes the round trip
        through memory.  Same arithmetic either way."""
        self.system, self.backend = system, system.backend
        self.fused = fused
        self.errest = bool(errest and self.bhat)
        self.nregs = 4 if self.errest else 2

        if system.nrhs < self.nregs:
            raise ValueError(f'{type(self).__name__} needs {self.nregs} '
                             'register banks')

        self.tcurr, self.idxcurr, self.nsteps = tstart, 0, 0
        self._regidx = list(range(self.nregs))

        self.backend.pointwise.register(
            'pyfr.integrators.explicit.kernels.rkvdh2'
        )

        self.c = [0.0] + [sum(self.b[:i]) + ai for i, ai in enumerate(self.a)]
        self.e = [b - bh for b, bh in zip(self.b, self.bhat)]
        self.nstages = len(self.c)
        self._kerns = {}

    def _stage_kerns(self, stage, r1, r2, *rs):
        key = (stage, r1, r2, *rs)

        if key not in self._kerns:
            tplargs = {
                'a': self.a, 'b': self.b, 'e': self.e, 'stage': stage,
                'nstages': self.nstages, 'nvars': self.system.nvars,
                'errest': bool(rs)
            }
            names = ('r1', 'r2', 'rold', 'rerr')

            self._kerns[key] = [
                self.backend.kernel(
                    'rkvdh2', tplargs=tplargs, dims=[shp[0], shp[2]],
                    **{n: em[r] for n, r in zip(names, key[1:])}
                )
                for shp, em in zip(self.system.ele_shapes.values(),
                                   self.system.ele_banks)
            ]

        return self._kerns[key]

    def _step(self, t, dt):
        r1 = self.idxcurr
        r2, *rs = sorted(set(self._regidx) - {r1})

        for i, ci in enumerate(self.c):
            kerns = self._stage_kerns(i, r1, r2, *rs)
            uin = r2 if i > 0 else r1

            if self.fused:
                self.system.rhs(t + ci*dt, uin, r2,
                                post=((i, r1, r2, *rs), kerns), dt=dt)
            else:
                self.system.rhs(t + ci*dt, uin, r2)
                for k in kerns:
                    k.bind(dt=dt)
                self.backend.run_kernels(kerns)

            r1, r2 = r2, r1

        return (r2, *rs) if rs else r2

    def advance(self, nsteps, dt):
        """Fixed step size (``controller = none``)."""
        for _ in range(nsteps):
            ret = self._step(self.tcurr, dt)
            self.idxcurr = ret[0] if self.errest else ret
            self.tcurr += dt
            self.nsteps += 1

    @property
    def soln(self):
        return self.system.ele_scal_upts(self.idxcurr)


class RK45Stepper(RKVdH2RStepper):
    # RK4(3)5[2R+]C of Kennedy, Carpenter & Lewis, Appl. Numer. Math. 35
    # (2000), table 8 -- the rationals PyFR's rk45 uses (steppers.py:218)
    stepper_order = 4

    a = [970286171893/4311952581923, 6584761158862/12103376702013,
         2251764453980/15575788980749, 26877169314380/34165994151039]

    b = [1153189308089/22510343858157, 1772645290293/4653164025191,
         -1672844663538/4480602732383, 2114624349019/3568978502595,
         5198255086312/14908931495163]

    bhat = [1016888040809/7410784769900, 11231460423587/58533540763752,
            -1563879915014/6823010717585, 606302364029/971179775848,
            1097981568119/3980877426909]


class NoneController:
    """Fixed nominal step size (``controller = none``,
    ``pyfr/integrators/explicit/controllers.py:46-63``); like the reference
    the last ``dt-lookahead`` steps before a target time are equalised."""

    sect = 'solver-time-integrator'

    def __init__(self, stepper, cfg):
        self.stepper, self.system = stepper, stepper.system
        self.backend = stepper.backend

        self.dt = cfg.getfloat(self.sect, 'dt')
        self.dtmin = cfg.getfloat(self.sect, 'dt-min', 1e-12)
        self._dt_lookahead = cfg.getint(self.sect, 'dt-lookahead', 10)

        self.nacptsteps = self.nrjctsteps = 0
        self.stepinfo = []
        self._tcomp = 0.0

    @property
    def tcurr(self):
        return self.stepper.tcurr

    def _advance_time(self, dt):
        # Compensated summation, as pyfr/integrators/base.py:282-287
        st = self.stepper
        y = dt - self._tcomp
        t = st.tcurr + y
        self._tcomp = (t - st.tcurr) - y
        st.tcurr = t

    def _clamp_dt(self, dt_want, t):
        remaining = t - self.tcurr
        nsteps = -(-remaining // dt_want)

        if nsteps > self._dt_lookahead:
            return dt_want
        else:
            return max(remaining / nsteps, self.dtmin)

    def _accept(self, dt, idxcurr, err=None):
        self._advance_time(dt)
        self.stepper.idxcurr = idxcurr
        self.stepper.nsteps += 1
        self.nacptsteps += 1
        self.stepinfo.append((dt, 'accept', err))

    def advance_to(self, t):
        if t < self.tcurr:
            raise ValueError('Advance time is in the past')

        st = self.stepper

        while self.tcurr < t:
            dt = self._clamp_dt(self.dt, t)
            ret = st._step(st.tcurr, dt)
            self._accept(dt, ret[0] if isinstance(ret, tuple) else ret)


class CFLController(NoneController):
    """``controller = cfl`` (pyfr/integrators/controllers.py:7-42): every
    ``cfl-nsteps`` accepted steps the step size is reset to
    ``cfl / (lambda_max (2p + 1))`` from the largest wave speed of the
    current solution.  The system must have been built with
    ``needs_cfl=True``."""

    def __init__(self, stepper, cfg, allreduce=None):
        super().__init__(stepper, cfg)
        self._allreduce = allreduce or (lambda x, op: x)

        self._order = cfg.getint('solver', 'order')
        self._cfl = cfg.getfloat(self.sect, 'cfl')
        self.dtmax = cfg.getfloat(self.sect, 'dt-max', 1e2)
        self._cfl_nsteps = cfg.getint(self.sect, 'cfl-nsteps', 1)

    def _compute_dt_cfl(self, uinbank):
        local_max = self.system.compute_max_wavespeed(uinbank)
        global_max = self._allreduce(local_max, 'max')
        return self._cfl / (global_max*(2*self._order + 1))

    def advance_to(self, t):
        if t < self.tcurr:
            raise ValueError('Advance time is in the past')

        st = self.stepper

        while self.tcurr < t:
            if self.nacptsteps % self._cfl_nsteps == 0:
                self.dt = self._compute_dt_cfl(st.idxcurr)

            dt = self._clamp_dt(min(self.dt, self.dtmax), t)
            ret = st._step(st.tcurr, dt)
            self._accept(dt, ret[0] if isinstance(ret, tuple) else ret)


class PIController(NoneController):
    """Adaptive step size for a stepper with an embedded error estimate.

    ``cfg`` supplies ``[solver-time-integrator]`` ``atol``, ``rtol``
    (or ``atol-<var>``), ``errest-norm``, ``safety-fact``, ``max-fact``,
    ``min-fact``, ``pi-alpha``, ``pi-beta``, ``dt``, ``dt-max``, ``dt-min``,
    ``dt-lookahead`` with the reference's defaults.  ``allreduce(x, op)``
    (op 'sum' or 'max') combines the ranks' error norms and DoF counts; the
    default is a single rank."""

    def __init__(self, stepper, cfg, convars, allreduce=None):
        if not stepper.errest:
            raise TypeError('Incompatible stepper/controller combination')

        super().__init__(stepper, cfg)
        be = self.backend
        self._allreduce = allreduce or (lambda x, op: x)

        f = lambda k, d=None: cfg.getfloat(self.sect, k, d)
        eps = float(np.finfo(be.fpdtype).eps)

        self.dtmax = f('dt-max', 1e2)

        self._rtol = f('rtol')
        if self._rtol < 10*eps:
            raise ValueError('Relative tolerance too small')

        has = [cfg.hasopt(self.sect, f'atol-{v}') for v in convars]
        if any(has) and not all(has):
            raise ValueError('Missing atol for some variables')

        if all(has):
            self._atols = tuple(f(f'atol-{v}') for v in convars)
        else:
            self._atols = (f('atol'),)*len(convars)

        if any(a < 10*eps for a in self._atols):
            raise ValueError('Absolute tolerance too small')

        self._norm = cfg.get(self.sect, 'errest-norm', 'l2')
        if self._norm not in {'l2', 'uniform'}:
            raise ValueError('Invalid error norm')

        self._saffac = f('safety-fact', 0.8)
        self._maxfac = f('max-fact', 1.1)
        self._minfac = f('min-fact', 0.9)
        if not self._minfac < 1 <= self._maxfac:
            raise ValueError('Invalid max-fact, min-fact')

        self._alpha, self._beta = f('pi-alpha', 0.58), f('pi-beta', 0.42)
        self._errprev = 1.0

        self.gndofs = self._allreduce(
            sum(int(np.prod(s)) for s in self.system.ele_shapes.values()),
            'sum'
        )

        self._ekerns = {}

    def _errest(self, rcurr, rerr):
        if (rcurr, rerr) not in self._ekerns:
            expr = 'err / (atol + rtol*fabs(curr))'
            if self._norm == 'uniform':
                expr, rop = f'fabs({expr})', 'max'
            else:
                expr, rop = f'({expr})*({expr})', 'sum'

            kerns = [
                self.backend.kernel('reduction', rop, [expr],
                                    {'curr': em[rcurr], 'err': em[rerr]},
                                    svars=['rtol'],
                                    pvars={'atol': self._atols})
                for em in self.system.ele_banks
            ]
            self.backend.commit()
            for k in kerns:
                k.bind(self._rtol)

            self._ekerns[rcurr, rerr] = kerns

        kerns = self._ekerns[rcurr, rerr]
        self.backend.run_kernels(kerns, wait=True)

        if self._norm == 'l2':
            err = self._allreduce(sum(float(k.retval[0]) for k in kerns),
                                  'sum')
            err = math.sqrt(err / self.gndofs)
        else:
            err = self._allreduce(max(float(k.retval[0]) for k in kerns),
                                  'max')

        return err if not math.isnan(err) else 100

    def advance_to(self, t):
        if t < self.tcurr:
            raise ValueError('Advance time is in the past')

        st = self.stepper
        expa, expb = self._alpha/st.stepper_order, self._beta/st.stepper_order

        while self.tcurr < t:
            dt = self._clamp_dt(min(self.dt, self.dtmax), t)

            icurr, iprev, ierr = st._step(st.tcurr, dt)
            err = self._errest(icurr, ierr)

            fac = err**-expa * self._errprev**expb
            fac = min(self._maxfac, max(self._minfac, self._saffac*fac))
            self.dt = fac*dt

            if err < 1.0:
                self._errprev = err
                self._accept(dt, icurr, err)
            else:
                if dt <= self.dtmin:
                    raise RuntimeError('Minimum sized time step rejected')

                st.idxcurr = iprev
                self.nrjctsteps += 1
                self.stepinfo.append((dt, 'reject', err))


def compile_expr(expr, privars, ndims):
    subs = {v: f'pri[{i}]' for i, v in enumerate(privars)}
    for i, v in enumerate(privars):
        for j, d in enumerate('xyz'[:ndims]):
            subs[f'grad_{v}_{d}'] = f'grad_pri[{i}][{j}]'
    for d, c in enumerate('xyz'[:ndims]):
        subs[c] = f'ploc[{d}]'

    p = '|'.join(re.escape(k) for k in sorted(subs, key=len, reverse=True))
    return re.sub(rf'\b({p})\b', lambda m: subs[m[1]], expr)


class FieldIntegrator:
    """Volume integrals of expressions, summed on the device per element."""

    def __init__(self, system, cfg, exprs, reduceop='sum'):
        self.system, self.backend = system, system.backend
        be = self.backend

        _, _, privars, ndims, nvars, _, _ = system.ele_quad[0]
        self.nexprs, self.reduceop = len(exprs), reduceop
        self.has_grads = bool(re.search(r'\bgrad_', ' '.join(exprs)))
        self.has_ploc = bool(re.search(r'\b[xyz]\b', ' '.join(exprs)))

        be.pointwise.register('pyfr.plugins.kernels.fieldeval')

        self._tplargs = {
            'ndims': ndims, 'nvars': nvars, 'nexprs': self.nexprs,
            'exprs': [compile_expr(e, privars, ndims) for e in exprs],
            'reduceop': reduceop, 'c': cfg.items_as('constants', float),
            'has_grads': self.has_grads, 'use_views': False,
            'has_wts': reduceop == 'sum',
            'eos_mod': 'pyfr.solvers.euler.kernels.eos'
        }

        self._edata = []
        for i, (wts, rcpdjac, _, _, _, nupts, neles) in \
                enumerate(system.ele_quad):
            w = (be.const_matrix(wts[:, None]/rcpdjac, tags={'align'})
                 if reduceop == 'sum' else None)
            ploc = (be.const_matrix(system.ele_ploc_upts[i](),
                                    tags={'align'})
                    if self.has_ploc else None)
            out = be.matrix((self.nexprs, neles), tags={'align'})
            self._edata.append((w, out, nupts, neles, ploc))

        be.commit()
        self._kerns = {}

    def total_volume(self):
        return sum(float((wts[:, None]/rcpdjac).sum())
                   for wts, rcpdjac, *_ in self.system.ele_quad)

    def kernels(self, uidx):
        be, sysm = self.backend, self.system

        if uidx not in self._kerns:
            self._kerns[uidx] = [
                be.pointwise.fieldeval(
                    tplargs=self._tplargs, dims=[nupts, neles],
                    u=sysm.ele_banks[i][uidx], out=out,
                    **({'gradu': sysm.eles_vect_upts[i]}
                       if self.has_grads else {}),
                    **({'wts': w} if w is not None else {}),
                    **({'ploc': ploc} if ploc is not None else {})
                )
                for i, (w, out, nupts, neles, ploc) in enumerate(self._edata)
            ]

        return self._kerns[uidx]

    def __call__(self, t, uidx):
        be, sysm = self.backend, self.system

        if self.has_grads:
            sysm.compute_grads(t, uidx)

        for k in self.kernels(uidx):
            if hasattr(k, 'bind'):
                k.bind(t=t)
        be.run_kernels(self.kernels(uidx))

        ident = {'sum': 0.0, 'min': np.inf, 'max': -np.inf}[self.reduceop]
        res = np.full(self.nexprs, ident)
        for w, out, *_ in self._edata:
            o = out.get()
            if self.reduceop == 'sum':
                res += o.sum(axis=1)
            elif self.reduceop == 'max':
                res = np.maximum(res, o.max(axis=1))
            else:
                res = np.minimum(res, o.min(axis=1))

        return res


# The documented Taylor-Green diagnostics
# (doc/src/plugins/soln-plugin-integrate.rst:64-70)
TGV_EXPRS = [
    '0.5*rho*(u*u + v*v + w*w)',
    '0.5*rho*((grad_w_y - grad_v_z)*(grad_w_y - grad_v_z) + '
    '(grad_u_z - grad_w_x)*(grad_u_z - grad_w_x) + '
    '(grad_v_x - grad_u_y)*(grad_v_x - grad_u_y))',
]
