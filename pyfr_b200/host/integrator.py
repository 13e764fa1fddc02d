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

Everything numerical is a backend kernel; this file only sequences them.
"""

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

    def step(self, dt):
        add, rhs, t = self._add, self.system.rhs, self.tcurr
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

        self.idxcurr = r1
        self.tcurr += dt
        self.nsteps += 1

        return r1

    def advance(self, nsteps, dt):
        for _ in range(nsteps):
            self.step(dt)

    @property
    def soln(self):
        return self.system.ele_scal_upts(self.idxcurr)


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

    def __init__(self, system, cfg, exprs):
        self.system, self.backend = system, system.backend
        be = self.backend

        _, _, privars, ndims, nvars, _, _ = system.ele_quad[0]
        self.nexprs = len(exprs)
        self.has_grads = bool(re.search(r'\bgrad_', ' '.join(exprs)))

        if re.search(r'\b[xyz]\b', ' '.join(exprs)):
            raise NotImplementedError('coordinate-dependent integrands')

        be.pointwise.register('pyfr.plugins.kernels.fieldeval')

        self._tplargs = {
            'ndims': ndims, 'nvars': nvars, 'nexprs': self.nexprs,
            'exprs': [compile_expr(e, privars, ndims) for e in exprs],
            'reduceop': 'sum', 'c': cfg.items_as('constants', float),
            'has_grads': self.has_grads, 'use_views': False,
            'has_wts': True,
            'eos_mod': 'pyfr.solvers.euler.kernels.eos'
        }

        self._edata = []
        for wts, rcpdjac, _, _, _, nupts, neles in system.ele_quad:
            w = be.const_matrix(wts[:, None]/rcpdjac, tags={'align'})
            out = be.matrix((self.nexprs, neles), tags={'align'})
            self._edata.append((w, out, nupts, neles))

        be.commit()
        self._kerns = {}

    def total_volume(self):
        return sum(float(w.get().sum()) for w, *_ in self._edata)

    def __call__(self, t, uidx):
        be, sysm = self.backend, self.system

        if self.has_grads:
            sysm.compute_grads(t, uidx)

        if uidx not in self._kerns:
            self._kerns[uidx] = [
                be.pointwise.fieldeval(
                    tplargs=self._tplargs, dims=[nupts, neles],
                    u=sysm.ele_banks[i][uidx], out=out, wts=w,
                    **({'gradu': sysm.eles_vect_upts[i]}
                       if self.has_grads else {})
                )
                for i, (w, out, nupts, neles) in enumerate(self._edata)
            ]

        for k in self._kerns[uidx]:
            if hasattr(k, 'bind'):
                k.bind(t=t)
        be.run_kernels(self._kerns[uidx])

        res = np.zeros(self.nexprs)
        for w, out, *_ in self._edata:
            res += out.get().sum(axis=1)

        return res


# The documented Taylor-Green diagnostics
# (doc/src/plugins/soln-plugin-integrate.rst:64-70)
TGV_EXPRS = [
    '0.5*rho*(u*u + v*v + w*w)',
    '0.5*rho*((grad_w_y - grad_v_z)*(grad_w_y - grad_v_z) + '
    '(grad_u_z - grad_w_x)*(grad_u_z - grad_w_x) + '
    '(grad_v_x - grad_u_y)*(grad_v_x - grad_u_y))',
]
