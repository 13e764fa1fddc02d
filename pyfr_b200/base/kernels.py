"""Kernel objects and the dependency graph they are scheduled through.

Mirrors the contract of the reference's ``pyfr/backends/base/provider.py``
(Kernel :10-25, NullKernel :28-29, meta kernels :32-73, NotSuitableError
:209-210) and of ``Graph`` in ``pyfr/backends/base/types.py:343-533``: the
same ``add/add_all/add_mpi_req(s)/group/commit/run`` surface and the same
scheduling policy (hard deps respected, kernels feeding an exchange send
hoisted first, insertion order otherwise, grouped kernels kept adjacent).
"""

import heapq


class NotSuitableError(Exception):
    pass


class Kernel:
    compound = False
    rtnames = ()

    def __init__(self, mats=(), views=(), misc=(), dt=float('nan')):
        self.mats = list(mats)
        self.views = list(views)
        self.misc = list(misc)
        self.dt = dt

    @property
    def retval(self):
        return None

    def run(self, *args):
        pass


class NullKernel(Kernel):
    pass


class MetaKernel(Kernel):
    def __init__(self, kernels):
        super().__init__()
        self.kernels = list(kernels)

        bindable = [k for k in self.kernels if hasattr(k, 'bind')]
        if bindable:
            self._bindable = bindable
            self.bind = self._bind_all

    def _bind_all(self, **kwargs):
        for k in self._bindable:
            k.bind(**kwargs)

    def run(self, *args):
        for k in self.kernels:
            k.run(*args)

    def add_to_graph(self, graph, deps):
        raise NotImplementedError


class OrderedMetaKernel(MetaKernel):
    pass


class UnorderedMetaKernel(MetaKernel):
    def __init__(self, kernels, splits=None):
        super().__init__(kernels)

        if splits is not None:
            self.splits = list(splits)
            self.compound = True

            if len(self.splits) != len(self.kernels) - 1:
                raise ValueError('Invalid split points')


class Graph:
    def __init__(self, backend):
        self.backend = backend
        self.committed = False

        self._order = []        # kernels in insertion order
        self.kdeps = {}         # hard dependencies
        self.kpdeps = {}        # pseudo (ordering-hint) dependencies
        self._xreqs = []        # (request, deps) exchange requests
        self._groups = []       # (kernels, substitutions)

        self.knodes = {}
        self.depk = set()
        self.mpi_reqs = []
        self.mpi_req_deps = []
        self.mpi_root_reqs = []

    def _check_open(self):
        if self.committed:
            raise RuntimeError('Can not modify a committed graph')

    def add(self, kern, deps=[], pdeps=[]):
        self._check_open()

        if kern in self.kdeps:
            raise RuntimeError('Can only add a kernel to a graph once')

        self._order.append(kern)
        self.kdeps[kern] = list(deps)
        self.kpdeps[kern] = list(pdeps)

    def add_all(self, kerns, deps=[], pdeps=[]):
        for k in kerns:
            self.add(k, deps, pdeps)

    def add_mpi_req(self, req, deps=[]):
        self._check_open()
        self._xreqs.append((req, list(deps)))

    def add_mpi_reqs(self, reqs, deps=[]):
        for r in reqs:
            self.add_mpi_req(r, deps)

    def group(self, kerns, subs=[]):
        self._check_open()
        self._groups.append((list(kerns), list(subs)))

    def _schedule(self):
        # Collapse every group onto its first member
        head = {}
        members = {}
        for kerns, _ in self._groups:
            members[kerns[0]] = kerns
            for k in kerns:
                head[k] = kerns[0]

        sup = lambda k: head.get(k, k)
        nodes = [k for k in self._order if sup(k) is k]
        pos = {k: i for i, k in enumerate(self._order)}
        rank = {n: min(pos[m] for m in members.get(n, [n])) for n in nodes}

        # Hard-dependency edges between super nodes
        preds = {n: set() for n in nodes}
        for n in nodes:
            for m in members.get(n, [n]):
                for d in self.kdeps[m]:
                    if sup(d) is not n:
                        preds[n].add(sup(d))

        succs = {n: [] for n in nodes}
        for n, ps in preds.items():
            for p in ps:
                succs[p].append(n)

        # Everything an exchange send transitively waits on is urgent
        urgent = set()
        todo = [sup(d) for _, deps in self._xreqs for d in deps]
        while todo:
            n = todo.pop()
            if n not in urgent:
                urgent.add(n)
                todo.extend(preds[n])

        key = lambda n: (n not in urgent, rank[n])
        indeg = {n: len(preds[n]) for n in nodes}
        ready = [(key(n), id(n), n) for n in nodes if not indeg[n]]
        heapq.heapify(ready)

        out = []
        while ready:
            _, _, n = heapq.heappop(ready)
            out.extend(members.get(n, [n]))

            for s in succs[n]:
                indeg[s] -= 1
                if not indeg[s]:
                    heapq.heappush(ready, (key(s), id(s), s))

        if len(out) != len(self._order):
            raise RuntimeError('Cyclic kernel dependencies')

        return out

    # Hooks for concrete backends
    def _add_mpi_req(self, req, deps):
        self.mpi_reqs.append(req)
        self.mpi_req_deps.append(deps)

    def _group(self, kerns, subs):
        pass

    def _commit(self):
        pass

    def commit(self):
        self.committed = True

        sched = self._schedule()
        where = {k: i for i, k in enumerate(sched)}

        # Sends fire after the last kernel they depend on
        sends_at = {}
        for req, deps in self._xreqs:
            if deps:
                at = max(where[d] for d in deps)
                sends_at.setdefault(at, []).append((req, deps))
            else:
                self.mpi_root_reqs.append(req)
                self._add_mpi_req(req, deps)

        for i, kern in enumerate(sched):
            alld = self.kdeps[kern] + self.kpdeps[kern]
            live = [d for d in alld if d in self.knodes]
            self.knodes[kern] = kern.add_to_graph(
                self, [self.knodes[d] for d in live]
            )
            self.depk.update(live)

            for req, deps in sends_at.get(i, []):
                self._add_mpi_req(req, deps)

        for kerns, subs in self._groups:
            self._group(kerns, subs)

        self.sched = sched
        self._commit()

    def run(self, *args):
        raise NotImplementedError

    def get_wait_times(self):
        return []
