"""TEST ORACLE / CPU BASELINE -- the C11 + OpenMP backend (oracle/crhs).

Test infrastructure only (see oracle/physics.py for who may import it).
It is the NumPy oracle backend with the hot kernels replaced by the C
restatement of the reference's CPU design in ``oracle/crhs/crhs.c``:
operator multiplies, ``tflux``/``gradcoru``/``negdivconf`` on linear
elements, the interface kernels and ``pack``.  Anything the C file does not
cover raises ``NotSuitableError`` and falls through to the NumPy provider,
exactly as providers are tried in turn by ``BaseBackend.kernel``
(pyfr/backends/base/backend.py:188-216).

Like the reference's OpenMP backend (``blocks = True``) it honours
``Graph.group``: the kernels of a group run block by block inside one
OpenMP loop with the group's private temporaries in thread-local scratch
(pyfr/backends/openmp/types.py:136-186).
"""

import ctypes as ct
import os
import subprocess

import numpy as np

from oracle.npbackend import _ViewRef, make_backend

_here = os.path.dirname(os.path.abspath(__file__))


def build():
    subprocess.run(['make', '-s', '-C', _here], check=True,
                   capture_output=True)


class KDesc(ct.Structure):
    _fields_ = [
        ('kind', ct.c_int), ('ndims', ct.c_int), ('nvars', ct.c_int),
        ('npts', ct.c_int), ('ld', ct.c_int), ('ksoa', ct.c_int),
        ('csub', ct.c_int), ('viscous', ct.c_int), ('nverts', ct.c_int),
        ('neles', ct.c_int),
        ('p', ct.c_void_p*4), ('bs', ct.c_long*4), ('slot', ct.c_int*4),
        ('off', ct.c_long*4),
        ('M', ct.c_int), ('rowptr', ct.c_void_p), ('cols', ct.c_void_p),
        ('vals', ct.c_void_p), ('beta', ct.c_double),
        ('gamma', ct.c_double), ('mu', ct.c_double),
        ('gamma_pr', ct.c_double), ('pts', ct.c_void_p),
    ]


class CfluxArgs(ct.Structure):
    _fields_ = [
        ('ndims', ct.c_int), ('nvars', ct.c_int), ('ksoa', ct.c_int),
        ('rsolver', ct.c_int), ('viscous', ct.c_int), ('mpi', ct.c_int),
        ('gamma', ct.c_double), ('mu', ct.c_double),
        ('gamma_pr', ct.c_double), ('beta', ct.c_double),
        ('tau', ct.c_double), ('n', ct.c_long), ('base', ct.c_void_p),
        ('gbase', ct.c_void_p),
        ('ul_map', ct.c_void_p), ('ur_map', ct.c_void_p),
        ('ur_mpi', ct.c_void_p), ('gl_map', ct.c_void_p),
        ('gl_str', ct.c_void_p), ('gr_map', ct.c_void_p),
        ('gr_str', ct.c_void_p), ('gr_mpi', ct.c_void_p),
        ('nl', ct.c_void_p), ('nl_ld', ct.c_long),
    ]


class ConuArgs(ct.Structure):
    _fields_ = [
        ('nvars', ct.c_int), ('ksoa', ct.c_int), ('mpi', ct.c_int),
        ('beta', ct.c_double), ('n', ct.c_long), ('base', ct.c_void_p),
        ('obase', ct.c_void_p), ('li_map', ct.c_void_p), ('ri_map', ct.c_void_p),
        ('lo_map', ct.c_void_p), ('ro_map', ct.c_void_p),
        ('ri_mpi', ct.c_void_p),
    ]


K_MUL, K_TFLUX, K_GRADCORU, K_NEGDIVCONF = range(4)
_argpos = {K_MUL: {'b': 0, 'out': 1}, K_TFLUX: {'u': 0, 'f': 1, 'verts': 2},
           K_GRADCORU: {'gradu': 0, 'verts': 1},
           K_NEGDIVCONF: {'tdivtconf': 0, 'rcpdjac': 1}}


def _ptr(a):
    return a.ctypes.data


def _mptr(m):
    """Address of a matrix/slice window inside its extent."""
    return m.basedata.ctypes.data + m.offset


def _root(m):
    return getattr(m, 'parent', m)


def make_cbackend(base, fast=False, name='oracle-c', nthreads=None):
    """``nthreads``: OpenMP threads to use (default: what the environment
    says; launchers such as torchrun export OMP_NUM_THREADS=1, so the
    benchmark passes the host's core count explicitly)."""
    build()
    lib = ct.CDLL(os.path.join(_here, '_build',
                               'libcrhs_fast.so' if fast else 'libcrhs.so'))
    lib.crhs_num_threads.restype = ct.c_int
    if nthreads:
        lib.crhs_set_num_threads(ct.c_int(int(nthreads)))

    NP = make_backend(base, name=name)
    NPKernel = NP.kernel_cls

    class BlockKernel(NPKernel):
        """An element kernel the block runner can execute."""

        def __init__(self, be, desc, nblocks, mats, keep):
            NPKernel.__init__(self, None)
            self.be, self.desc, self.nblocks = be, desc, nblocks
            self.argmats, self.keep = mats, keep

        def run(self, *args):
            arr = (KDesc*1)(self.desc)
            lib.crhs_run_blocks(1, arr, self.nblocks, 0, None)

    class GroupKernel(NPKernel):
        def __init__(self, kerns, subs):
            NPKernel.__init__(self, None)
            self.kerns = kerns

            descs = []
            for k in kerns:
                d = KDesc()
                ct.memmove(ct.byref(d), ct.byref(k.desc), ct.sizeof(KDesc))
                descs.append(d)

            # Thread-local scratch substitution
            slots = []
            for sub in subs:
                roots = {id(_root(k.argmats[a])) for k, a in sub}
                if len(roots) != 1:
                    continue

                r = _root(sub[0][0].argmats[sub[0][1]])
                slot = len(slots)
                slots.append(r.nrow*r.leaddim)

                for k, a in sub:
                    d = descs[kerns.index(k)]
                    pos = _argpos[d.kind][a]
                    m = k.argmats[a]
                    d.slot[pos] = slot
                    d.off[pos] = getattr(m, 'ra', 0)*m.leaddim

            self.descs = (KDesc*len(descs))(*descs)
            self.slots = (ct.c_long*max(len(slots), 1))(*slots)
            self.nslots = len(slots)
            self.nblocks = kerns[0].nblocks

        def run(self, *args):
            lib.crhs_run_blocks(len(self.kerns), self.descs, self.nblocks,
                                self.nslots, self.slots)

    class FnKernel(NPKernel):
        def __init__(self, fn, args, keep):
            NPKernel.__init__(self, None)
            self.fn, self.args, self.keep = fn, args, keep

        def run(self, *a):
            self.fn(ct.byref(self.args))

    def block_desc(be, kind, mats, **kw):
        """Fill the common part of a descriptor from its matrix args."""
        d = KDesc()
        d.kind = kind
        for a, m in mats.items():
            pos = _argpos[kind][a]
            d.p[pos] = _mptr(m)
            d.bs[pos] = m.blocksz
            d.slot[pos] = -1
        for pos in range(4):
            if not d.p[pos]:
                d.slot[pos] = -1
        d.ksoa, d.csub = be.soasz, be.csubsz
        for k, v in kw.items():
            setattr(d, k, v)
        return d

    def blocked(m):
        return (m.backend.blocks and m.dtype == np.float64 and
                not ({'xchg', 'noblock'} & set(m.tags)))

    class CBlasProvider:
        def __init__(self, backend):
            self.backend = backend

        def mul(self, a, b, out, alpha=1.0, beta=0.0):
            be = self.backend

            if a.nrow != out.nrow or a.ncol != b.nrow or b.ncol != out.ncol:
                raise ValueError('Incompatible matrices for out = a*b')
            if not (blocked(b) and blocked(out)) or \
               b.leaddim != out.leaddim or be.extended_mul:
                raise base.NotSuitableError('not a blocked fp64 multiply')

            A = alpha*a.get()
            rowptr = np.zeros(A.shape[0] + 1, dtype=np.int32)
            cols, vals = [], []
            for r, row in enumerate(A):
                nz = np.flatnonzero(row)
                cols += list(nz)
                vals += list(row[nz])
                rowptr[r + 1] = len(cols)
            cols = np.array(cols, dtype=np.int32)
            vals = np.array(vals, dtype=np.float64)

            d = block_desc(be, K_MUL, {'b': b, 'out': out}, M=A.shape[0],
                           ld=b.leaddim, rowptr=_ptr(rowptr),
                           cols=_ptr(cols), vals=_ptr(vals), beta=beta,
                           neles=b.nblocks*be.csubsz)
            return BlockKernel(be, d, b.nblocks, {'b': b, 'out': out},
                               [rowptr, cols, vals, a, b, out])

        def copy(self, dst, src):
            if dst.traits != src.traits:
                raise ValueError('Incompatible matrix types')
            if not (blocked(dst) and blocked(src)):
                raise base.NotSuitableError('not blocked')

            n = dst.nrow*dst.leaddim

            class CopyKernel(NPKernel):
                def run(self, *a):
                    lib.crhs_copy_rows(
                        ct.c_int(dst.nblocks), ct.c_long(n),
                        ct.c_void_p(_mptr(dst)), ct.c_long(dst.blocksz),
                        ct.c_void_p(_mptr(src)), ct.c_long(src.blocksz))

            return CopyKernel(None)

        def pack(self, xv):
            be = self.backend
            v, xm = xv.view, xv.xchgmat
            mp = np.ascontiguousarray(v.mapping.get()[0], dtype=np.int32)
            st = (np.ascontiguousarray(v.rstrides.get()[0], dtype=np.int32)
                  if v.rstrides is not None else None)
            basep = v.basedata.ctypes.data

            class PackKernel(NPKernel):
                def run(self, *a):
                    lib.crhs_pack(
                        ct.c_long(xv.n), ct.c_int(xv.nvrow),
                        ct.c_int(xv.nvcol), ct.c_int(be.soasz),
                        ct.c_void_p(basep), ct.c_void_p(_ptr(mp)),
                        ct.c_void_p(_ptr(st) if st is not None else 0),
                        ct.c_void_p(_mptr(xm)))

            return PackKernel(None)

    def _imap(v):
        v = getattr(v, 'view', v)
        return np.ascontiguousarray(v.mapping.get()[0], dtype=np.int32)

    def _istr(v):
        v = getattr(v, 'view', v)
        return np.ascontiguousarray(v.rstrides.get()[0], dtype=np.int32)

    def _vbase(v):
        v = getattr(v, 'view', v)
        return v.basedata.ctypes.data

    class CPointwise:
        """C versions of the pointwise kernels (linear elements)."""

        def __init__(self, backend):
            self.backend = backend

        def _phys(self, c):
            return dict(gamma=c['gamma'], mu=c.get('mu', 0.0),
                        gamma_pr=c['gamma']/c['Pr'] if 'Pr' in c else 0.0)

        def tflux(self, tplargs, dims, extrns={}, u=None, f=None,
                  gradu=None, smats=None, rcpdjac=None, verts=None,
                  upts=None, **kw):
            be = self.backend
            mod = be.pointwise._mods.get('tflux', '')
            viscous = 'navstokes' in mod

            if (tplargs['ktype'] != 'linear' or not blocked(u) or
                tplargs.get('visc_corr', 'none') != 'none' or
                tplargs.get('shock_capturing', 'none') != 'none'):
                raise base.NotSuitableError('C tflux: linear, unfused only')

            npts, neles = dims
            pts = np.ascontiguousarray(upts.get(), dtype=np.float64)
            d = block_desc(be, K_TFLUX, {'u': u, 'f': f, 'verts': verts},
                           ndims=tplargs['ndims'], nvars=tplargs['nvars'],
                           npts=npts, ld=u.leaddim, viscous=int(viscous),
                           nverts=tplargs['nverts'], neles=neles,
                           pts=_ptr(pts), **self._phys(tplargs['c']))
            return BlockKernel(be, d, -(-neles // be.csubsz),
                               {'u': u, 'f': f, 'verts': verts},
                               [pts, u, f, verts])

        def gradcoru(self, tplargs, dims, extrns={}, gradu=None, smats=None,
                     rcpdjac=None, verts=None, upts=None, **kw):
            be = self.backend
            if tplargs['ktype'] != 'linear' or not blocked(gradu):
                raise base.NotSuitableError('C gradcoru: linear only')

            npts, neles = dims
            pts = np.ascontiguousarray(upts.get(), dtype=np.float64)
            d = block_desc(be, K_GRADCORU, {'gradu': gradu, 'verts': verts},
                           ndims=tplargs['ndims'], nvars=tplargs['nvars'],
                           npts=npts, ld=gradu.leaddim,
                           nverts=tplargs['nverts'], neles=neles,
                           pts=_ptr(pts))
            return BlockKernel(be, d, -(-neles // be.csubsz),
                               {'gradu': gradu, 'verts': verts},
                               [pts, gradu, verts])

        def negdivconf(self, tplargs, dims, extrns={}, tdivtconf=None,
                       rcpdjac=None, ploc=None, u=None, **kw):
            be = self.backend
            if tplargs['src_macros'] or not blocked(tdivtconf):
                raise base.NotSuitableError('C negdivconf: no sources')

            npts, neles = dims
            d = block_desc(be, K_NEGDIVCONF,
                           {'tdivtconf': tdivtconf, 'rcpdjac': rcpdjac},
                           nvars=tplargs['nvars'], npts=npts,
                           ld=tdivtconf.leaddim, neles=neles)
            k = BlockKernel(be, d, -(-neles // be.csubsz),
                            {'tdivtconf': tdivtconf, 'rcpdjac': rcpdjac},
                            [tdivtconf, rcpdjac])
            k.rtnames = ('t',)
            k.bind = lambda **kw: None
            return k

        def _cflux(self, viscous, mpi, tplargs, dims, ul, ur, gradul, gradur,
                   nl):
            be = self.backend
            c = tplargs['c']
            if tplargs.get('visc_corr', 'none') != 'none' or \
               be.fpdtype != np.float64:
                raise base.NotSuitableError('C cflux: fp64, no Sutherland')

            a = CfluxArgs()
            a.ndims, a.nvars = tplargs['ndims'], tplargs['nvars']
            a.ksoa = be.soasz
            a.rsolver = {'rusanov': 0, 'hllc': 1}[tplargs['rsolver']]
            a.viscous, a.mpi = int(viscous), int(mpi)
            a.gamma = c['gamma']
            a.n = dims[0]
            a.base = _vbase(ul)
            keep = [ul, ur, gradul, gradur, nl]

            def put(name, arr):
                keep.append(arr)
                setattr(a, name, _ptr(arr))

            put('ul_map', _imap(ul))
            if mpi:
                a.ur_mpi = _mptr(ur)
            else:
                put('ur_map', _imap(ur))

            if viscous:
                a.mu, a.gamma_pr = c['mu'], c['gamma']/c['Pr']
                a.beta, a.tau = c['ldg-beta'], c['ldg-tau']
                a.gbase = _vbase(gradul)
                if a.beta != -0.5:
                    put('gl_map', _imap(gradul))
                    put('gl_str', _istr(gradul))
                if a.beta != 0.5:
                    if mpi:
                        a.gr_mpi = _mptr(gradur)
                    else:
                        put('gr_map', _imap(gradur))
                        put('gr_str', _istr(gradur))

            nlm = np.ascontiguousarray(nl.get(), dtype=np.float64)
            put('nl', nlm)
            a.nl_ld = nlm.shape[1]

            return FnKernel(lib.crhs_cflux, a, keep)

        def intcflux(self, tplargs, dims, extrns={}, ul=None, ur=None,
                     gradul=None, gradur=None, artvisc=None, nl=None, **kw):
            visc = 'navstokes' in self.backend.pointwise._mods['intcflux']
            return self._cflux(visc, False, tplargs, dims, ul, ur, gradul,
                               gradur, nl)

        def mpicflux(self, tplargs, dims, extrns={}, ul=None, ur=None,
                     gradul=None, gradur=None, artvisc=None, nl=None, **kw):
            visc = 'navstokes' in self.backend.pointwise._mods['mpicflux']
            return self._cflux(visc, True, tplargs, dims, ul, ur, gradul,
                               gradur, nl)

        def _conu(self, mpi, tplargs, dims, ulin, urin, ulout, urout):
            be = self.backend
            if be.fpdtype != np.float64:
                raise base.NotSuitableError('C conu: fp64')

            a = ConuArgs()
            a.nvars, a.ksoa, a.mpi = tplargs['nvars'], be.soasz, int(mpi)
            a.beta, a.n = tplargs['c']['ldg-beta'], dims[0]
            a.base = _vbase(ulin)
            a.obase = _vbase(ulout)
            keep = [ulin, urin, ulout, urout]

            def put(name, arr):
                keep.append(arr)
                setattr(a, name, _ptr(arr))

            put('li_map', _imap(ulin))
            put('lo_map', _imap(ulout))
            if mpi:
                a.ri_mpi = _mptr(urin)
            else:
                put('ri_map', _imap(urin))
                put('ro_map', _imap(urout))

            return FnKernel(lib.crhs_conu, a, keep)

        def intconu(self, tplargs, dims, extrns={}, ulin=None, urin=None,
                    ulout=None, urout=None, **kw):
            return self._conu(False, tplargs, dims, ulin, urin, ulout, urout)

        def mpiconu(self, tplargs, dims, extrns={}, ulin=None, urin=None,
                    ulout=None, **kw):
            return self._conu(True, tplargs, dims, ulin, urin, ulout, None)

    class CGraph(NP.graph_cls):
        def _group(self, kerns, subs):
            self._cgroups = getattr(self, '_cgroups', []) + [(kerns, subs)]

        def _commit(self):
            def leaves(k):
                if hasattr(k, 'kernels'):
                    return [l for c in k.kernels for l in leaves(c)]
                return [k]

            # Expand meta kernels, then merge each group's block kernels
            prog = []
            for what, obj in self.program:
                if what == 'kernel':
                    prog += [('kernel', l) for l in leaves(obj)]
                else:
                    prog.append((what, obj))

            for kerns, subs in getattr(self, '_cgroups', []):
                flat = [l for k in kerns for l in leaves(k)]
                if not all(isinstance(l, BlockKernel) for l in flat) or \
                   len({l.nblocks for l in flat}) != 1:
                    continue

                # Substitutions name (possibly meta) kernels: expand
                fsubs = [[(l, a) for k, a in sub for l in leaves(k)
                          if a in l.argmats] for sub in subs]
                gk = GroupKernel(flat, fsubs)

                ids, done, new = {id(l) for l in flat}, False, []
                for what, obj in prog:
                    if what == 'kernel' and id(obj) in ids:
                        if not done:
                            new.append(('kernel', gk))
                            done = True
                    else:
                        new.append((what, obj))
                prog = new

            self.program = prog

    class CBackend(NP):
        graph_cls = CGraph
        blocks = True

        def __init__(self, cfg):
            if not cfg.hasopt('backend-oracle', 'blocks'):
                cfg.set('backend-oracle', 'blocks', 1)
            super().__init__(cfg)

            self.cpoint = CPointwise(self)
            self._providers = [CBlasProvider(self), self.cpoint,
                               *self._providers]
            self.nthreads = lib.crhs_num_threads()

    CBackend.name = name
    return CBackend
