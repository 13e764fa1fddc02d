"""ctypes binding of ``libpyfr_b200.so`` (see ``include/pyfr_b200.h``).

Follows the reference's FFI idiom (``pyfr/ctypesutil.py:8-41``: a table of
``(restype, name, *argtypes)`` entries, integer status codes mapped to
exceptions by an ``errcheck`` hook) with one difference: the error text
comes from the library (``b200_last_error``) instead of a code table.

There is no CPU fallback.  ``DryRuntime`` exists so that kernels can be
*generated and compiled* on a machine without a GPU (``__graft_entry__.
build()`` and the CPU-side tests); every attempt to move data or launch
through it raises.
"""

import ctypes as ct
from ctypes import (POINTER, byref, c_char_p, c_float, c_int, c_size_t,
                    c_uint, c_void_p)
import os


class B200Error(RuntimeError):
    pass


class B200NoDevice(B200Error):
    pass


_here = os.path.dirname(os.path.abspath(__file__))
_libpath = os.environ.get('PYFR_B200_LIBRARY_PATH',
                          os.path.join(_here, 'libpyfr_b200.so'))

_vp, _vpp = c_void_p, POINTER(c_void_p)

_functions = [
    (c_int, 'b200_init', c_int),
    (c_int, 'b200_device_info', POINTER(c_int), POINTER(c_int),
     POINTER(c_int), POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t)),
    (c_int, 'b200_malloc', _vpp, c_size_t),
    (c_int, 'b200_free', _vp),
    (c_int, 'b200_malloc_host', _vpp, c_size_t),
    (c_int, 'b200_free_host', _vp),
    (c_int, 'b200_memset', _vp, c_int, c_size_t, _vp),
    (c_int, 'b200_memcpy', _vp, _vp, c_size_t),
    (c_int, 'b200_memcpy_async', _vp, _vp, c_size_t, _vp),
    (c_int, 'b200_memcpy2d_async', _vp, c_size_t, _vp, c_size_t, c_size_t,
     c_size_t, _vp),
    (c_int, 'b200_stream_create', _vpp),
    (c_int, 'b200_stream_create_priority', _vpp, c_int),
    (c_int, 'b200_stream_destroy', _vp),
    (c_int, 'b200_stream_sync', _vp),
    (c_int, 'b200_device_sync'),
    (c_int, 'b200_event_create', _vpp),
    (c_int, 'b200_event_destroy', _vp),
    (c_int, 'b200_event_record', _vp, _vp),
    (c_int, 'b200_event_sync', _vp),
    (c_int, 'b200_event_elapsed_ms', POINTER(c_float), _vp, _vp),
    (c_int, 'b200_stream_wait_event', _vp, _vp),
    (c_int, 'b200_nvrtc_compile', c_char_p, c_char_p, POINTER(c_char_p),
     c_int, _vpp, POINTER(c_size_t), POINTER(c_char_p)),
    (c_int, 'b200_buffer_free', _vp),
    (c_int, 'b200_module_load', _vpp, _vp),
    (c_int, 'b200_module_unload', _vp),
    (c_int, 'b200_module_get_function', _vpp, _vp, c_char_p),
    (c_int, 'b200_function_set_dynamic_smem', _vp, c_int),
    (c_int, 'b200_function_info', _vp, POINTER(c_int), POINTER(c_int),
     POINTER(c_int), POINTER(c_int)),
    (c_int, 'b200_launch', _vp, c_uint, c_uint, c_uint, c_uint, c_uint,
     c_uint, c_uint, _vp, _vpp),
    (c_int, 'b200_capture_begin', _vp),
    (c_int, 'b200_capture_end', _vp, _vpp),
    (c_int, 'b200_graph_launch', _vp, _vp),
    (c_int, 'b200_graph_destroy', _vp),
    (c_int, 'b200_nccl_unique_id', c_char_p),
    (c_int, 'b200_nccl_init', _vpp, c_int, c_int, c_char_p),
    (c_int, 'b200_nccl_destroy', _vp),
    (c_int, 'b200_nccl_group_start'),
    (c_int, 'b200_nccl_group_end'),
    (c_int, 'b200_nccl_send', _vp, _vp, c_size_t, c_int, c_int, _vp),
    (c_int, 'b200_nccl_recv', _vp, _vp, c_size_t, c_int, c_int, _vp),
    (c_int, 'b200_nccl_allreduce', _vp, _vp, _vp, c_size_t, c_int, c_int,
     _vp),
]

exported_symbols = [f[1] for f in _functions] + ['b200_last_error']


class Runtime:
    """The loaded library; attribute access yields checked functions with
    the ``b200_`` prefix dropped (``rt.malloc``, ``rt.launch`` ...)."""

    dry = False

    def __init__(self, path=_libpath):
        try:
            self._lib = lib = ct.CDLL(path)
        except OSError as e:
            raise B200Error(
                f'Unable to load {path}: {e}; build it with '
                '`python -c "import __graft_entry__ as g; g.build()"`'
            )

        lib.b200_last_error.restype = c_char_p
        self._raw = {}

        for restype, name, *argtypes in _functions:
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = restype, argtypes
            fn.errcheck = self._errcheck

            short = name.removeprefix('b200_')
            self._raw[short] = fn
            if not hasattr(type(self), short):
                setattr(self, short, fn)

    def _errcheck(self, status, fn, args):
        if status != 0:
            msg = self._lib.b200_last_error().decode(errors='replace')
            cls = B200NoDevice if 'NoDevice' in msg or 'Insufficient' in msg \
                else B200Error
            raise cls(f'{fn.__name__}: {msg}')

    # Conveniences over the raw entry points
    def new_ptr(self, fn, *args):
        p = c_void_p()
        fn(byref(p), *args)
        return p.value

    def end_capture(self, stream):
        g = c_void_p()
        self.capture_end(stream, byref(g))
        return g.value

    def elapsed_ms(self, start, stop):
        ms = c_float()
        self.event_elapsed_ms(byref(ms), start, stop)
        return ms.value

    def function_attrs(self, func):
        v = [c_int() for _ in range(4)]
        self.function_info(func, *map(byref, v))
        return dict(zip(('nregs', 'static_smem', 'local_bytes',
                         'max_threads'), (x.value for x in v)))

    def device_info(self):
        sm, maj, mnr = c_int(), c_int(), c_int()
        tot, free, smem = c_size_t(), c_size_t(), c_size_t()
        self._raw['device_info'](byref(sm), byref(maj), byref(mnr),
                                 byref(tot), byref(free), byref(smem))
        return dict(sm_count=sm.value, cc=(maj.value, mnr.value),
                    total_mem=tot.value, free_mem=free.value,
                    smem_optin=smem.value)

    def nvrtc(self, src, name, opts):
        arr = (c_char_p*len(opts))(*[o.encode() for o in opts])
        img, n, log = c_void_p(), c_size_t(), c_void_p()

        try:
            self.nvrtc_compile(src.encode(), name.encode(), arr, len(opts),
                               byref(img), byref(n),
                               ct.cast(byref(log), POINTER(c_char_p)))
            return ct.string_at(img.value, n.value)
        finally:
            # both buffers are malloc'ed by the library
            for buf in (img, log):
                if buf.value:
                    self.buffer_free(buf)


class DryRuntime:
    """Build-time stand-in used where no GPU exists: hands out fake device
    addresses so that layouts, argument lists and kernel *sources* can be
    produced and compiled with nvcc.  It cannot move data or launch."""

    dry = True

    def __init__(self):
        self._next = 0x7f0000000000

    def new_ptr(self, fn, *args):
        if fn == 'malloc':
            p = self._next
            self._next += -(-max(args[0], 1) // 512)*512
            return p
        return 1

    def __getattr__(self, name):
        if name in ('malloc', 'stream_create', 'event_create',
                    'stream_create_priority'):
            return name
        if name in ('free', 'stream_destroy', 'event_destroy',
                    'module_unload', 'graph_destroy', 'free_host'):
            return lambda *a: None

        def refuse(*a, **k):
            raise B200NoDevice(f'b200_{name}: no CUDA device (dry build-only '
                               'runtime); the B200 backend has no CPU path')
        return refuse

    def device_info(self):
        return dict(sm_count=148, cc=(10, 0), total_mem=0, free_mem=0,
                    smem_optin=232448)


def load_runtime(device=0, dry=False):
    if dry:
        return DryRuntime()

    rt = Runtime()
    rt.init(device)
    return rt
