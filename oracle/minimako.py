"""TEST ORACLE -- a minimal renderer for the reference's kernel templates.

The arithmetic of every pointwise kernel of the reference lives in Mako
templates (``pyfr/solvers/*/kernels/*.mako``) and Mako is not installable
offline, which is why ``oracle/physics.py`` *restates* that arithmetic.
This module closes the loop: it renders the reference's own template
files -- read from ``/root/reference`` at test time, never copied -- with
just enough of the Mako language for those files, so that the macros and
kernel bodies can be compiled as C and the restatement checked against
them (``tests/test_oracle_templates.py``).

What is interpreted here: ``${expr}``, ``% for/if/elif/else/end*`` lines,
``<% python %>`` blocks, ``<%include file=.../>``, ``<%pyfr:macro>``,
``<%pyfr:alias/>`` and ``<%pyfr:kernel>`` (the body is captured, the
argument attributes are returned as rendered text); ``<%inherit/>`` and
``<%namespace/>`` are dropped.  What is *not* reimplemented: the helper
functions templates call (``pyfr.dot``, ``pyfr.array``, ``pyfr.ndrange``,
``pyfr.expand`` with its local-variable renaming) -- those are the
reference's own ``pyfr/backends/base/makoutil.py``, imported and called.

Only usable where ``/root/reference`` exists.
"""

import inspect
import os
import re

from oracle import refharness as rh

_NAMES = r'pyfr:macro|pyfr:kernel|pyfr:alias|include|inherit|namespace'
_TAG = re.compile(
    rf'</%(?P<cname>{_NAMES})\s*>|'
    rf'<%(?P<name>{_NAMES})\b(?P<attrs>[^>]*?)(?P<self>/)?>|'
    rf'<%(?!{_NAMES})(?P<py>.*?)%>',
    re.S
)
_ATTR = re.compile(r'''(\w+)\s*=\s*(['"])(.*?)\2''', re.S)


def _find_expr_end(text, i):
    """Index of the ``}`` closing the ``${`` whose body starts at ``i``."""
    depth, quote = 1, None
    while i < len(text):
        ch = text[i]
        if quote:
            if ch == '\\':
                i += 1
            elif ch == quote:
                quote = None
        elif ch in '\'"':
            quote = ch
        elif ch == '{':
            depth += 1
        elif ch == '}':
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError('unterminated ${')


class _Gen:
    def __init__(self):
        self.lines, self.ind, self.n = [], 1, 0

    def emit(self, s):
        self.lines.append('    '*self.ind + s)

    def text(self, s):
        # literal text with ${...} substitutions
        i = 0
        while True:
            j = s.find('${', i)
            if j < 0:
                break
            if j > i:
                self.emit(f'_w({s[i:j]!r})')
            k = _find_expr_end(s, j + 2)
            expr = ' '.join(s[j + 2:k].split('\n'))
            self.emit(f'_w(str({expr.strip()}))')
            i = k + 1
        if i < len(s):
            self.emit(f'_w({s[i:]!r})')


def _attr_expr(val):
    """Python expression evaluating a tag attribute containing ``${}``."""
    parts, i = [], 0
    while True:
        j = val.find('${', i)
        if j < 0:
            break
        if j > i:
            parts.append(repr(val[i:j]))
        k = _find_expr_end(val, j + 2)
        parts.append(f'str({val[j + 2:k]})')
        i = k + 1
    if i < len(val) or not parts:
        parts.append(repr(val[i:]))
    return ' + '.join(parts)


def _compile(text):
    g = _Gen()
    pos = 0

    def chunk(s):
        # text between tags: split off the '%' control lines
        buf = []
        for line in s.splitlines(keepends=True):
            st = line.strip()
            if st.startswith('##'):
                continue
            if st.startswith('%') and not st.startswith('%%'):
                if buf:
                    g.text(''.join(buf))
                    buf = []
                stmt = st[1:].strip()
                kw = stmt.split()[0].rstrip(':') if stmt else ''
                if kw.startswith('end'):
                    g.ind -= 1
                elif kw in ('elif', 'else', 'except', 'finally'):
                    g.ind -= 1
                    g.emit(stmt)
                    g.ind += 1
                else:
                    g.emit(stmt)
                    g.ind += 1
                    g.emit('pass')
            else:
                buf.append(line)
        if buf:
            g.text(''.join(buf))

    stack = []
    for m in _TAG.finditer(text):
        chunk(text[pos:m.start()])
        pos = m.end()

        if m['py'] is not None:
            code = inspect.cleandoc(m['py']) if '\n' in m['py'] \
                else m['py'].strip()
            for l in code.splitlines():
                g.emit(l)
            continue

        name, attrs = m['name'], dict((a, v) for a, _, v in
                                      _ATTR.findall(m['attrs'] or ''))

        if m['cname']:
            kind, fn, a = stack.pop()
            if kind != m['cname']:
                raise ValueError(f'mismatched </%{m["cname"]}>')
            g.emit("return ''.join(_o)")
            g.ind -= 1
            if kind == 'pyfr:macro':
                g.emit(f'_macro({a["name"]!r}, {a.get("params", "")!r}, '
                       f'{a.get("externs", "")!r}, {fn})')
            else:
                ka = ', '.join(f'{k!r}: {_attr_expr(v)}' for k, v in a.items()
                               if k not in ('name', 'ndim'))
                g.emit(f'_ndim[{a["name"]!r}] = {a.get("ndim", "1")!r}')
                g.emit(f'_kernel({a["name"]!r}, {{{ka}}}, {fn})')
        elif name in ('inherit', 'namespace'):
            continue
        elif name == 'include':
            g.emit(f'_include({_attr_expr(attrs["file"])})')
        elif name == 'pyfr:alias':
            g.emit(f'_alias({attrs["name"]!r}, {attrs["func"]!r})')
        else:
            g.n += 1
            fn = f'_body{g.n}'
            pyargs = ''
            if name == 'pyfr:macro':
                ps = [p.strip() for p in attrs.get('params', '').split(',')]
                pyargs = ', '.join(p[3:] for p in ps if p.startswith('py:'))
                attrs['params'] = ', '.join(p for p in ps
                                            if not p.startswith('py:'))
            g.emit(f'def {fn}({pyargs}):')
            g.ind += 1
            g.emit('_o = []')
            g.emit('_w = _o.append')
            stack.append((name, fn, attrs))

    chunk(text[pos:])
    return 'def _render(_w):\n' + '\n'.join(g.lines or ['    pass'])


class Renderer:
    """Renders reference templates addressed by module path
    (``'pyfr.solvers.euler.kernels.flux'``) with the template arguments
    ``tplargs``; collects macros and kernel bodies."""

    def __init__(self, tplargs, extrns=(), generator=None, fpdtype=None,
                 ixdtype=None):
        """With ``generator`` (a kernel generator class of the reference,
        e.g. its OpenMP one) ``<%pyfr:kernel>`` tags are turned into
        complete kernels exactly as ``makoutil.kernel`` does; they land in
        ``self.sources`` / ``self.argspecs``.  ``extrns`` is then the
        mapping of external argument specs."""
        rh.install_stubs()
        import pyfr.backends.base.makoutil as mu
        import pyfr.util as util

        # the stubs stand in for mako.runtime: a captured body is simply
        # the string its function returns
        mu.capture = lambda ctx, fn, *a, **kw: fn(*a, **kw)

        self.mu = mu
        self.kernels, self._seen = {}, set()
        self.sources, self.argspecs = {}, {}
        self.ctx = {'_macros': {}, '_extrns': (dict(extrns) if generator else
                                               {e: None for e in extrns})}

        ctx = self.ctx

        class PyFR:
            pass

        ns = PyFR()
        for fn in ('dot', 'array', 'ndrange', 'carray', 'ilog2range',
                   'polyfit', 'expand'):
            setattr(ns, fn, (lambda f: lambda *a, **kw: f(ctx, *a, **kw))(
                getattr(mu, fn)))

        def macro(name, params, externs, body):
            if name in ctx['_macros']:
                return
            ps = [p.strip() for p in params.split(',') if p.strip()]
            es = [e.strip() for e in externs.split(',') if e.strip()]
            ctx['_macros'][name] = mu.Macro(ps, es, inspect.signature(body),
                                            body, id(body))

        def alias(name, func):
            ctx['_macros'][name] = ctx['_macros'][func]

        def kernel(name, attrs, body):
            text = body()
            self.kernels[name] = (attrs, text)

            if generator is not None:
                # pyfr/backends/base/makoutil.py, kernel()
                if any(a in ctx['_extrns'] for a in attrs):
                    raise ValueError(f'Duplicate argument in {name}')
                kern = generator(name, int(self._ndim[name]),
                                 dict(attrs, **ctx['_extrns']), text, fpdtype,
                                 ixdtype)
                self.argspecs[name] = kern.argspec()
                self.sources[name] = kern.render()

        self._ndim = {}

        import math
        self.ns = dict(tplargs, pyfr=ns, math=math, _macro=macro,
                       _alias=alias, _kernel=kernel, _include=self.include,
                       _ndim=self._ndim)

    def include(self, mod):
        if mod in self._seen:
            return
        self._seen.add(mod)

        path = os.path.join(rh.REFROOT, *mod.split('.')) + '.mako'
        with open(path) as f:
            code = _compile(f.read())

        scope = dict(self.ns)
        exec(code, scope)
        scope['_render'](lambda s: None)

    def expand(self, name, *args, **kw):
        """The C text of one macro invocation."""
        return self.mu.expand(self.ctx, name, *args, **kw)
