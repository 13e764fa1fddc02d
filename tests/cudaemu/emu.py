"""TEST INFRASTRUCTURE -- run the product's generated CUDA kernels on the
CPU (see emu.h).  ``install(monkeypatch)`` swaps the backend's runtime for
``EmuRuntime`` and the kernel compiler for a g++ build of the same source
text, so ``B200Backend`` executes end to end without a GPU: same host
code, same generators, same fusion decisions, same kernel text.  Only the
execution substrate differs (OS threads instead of CUDA threads, memcpy
instead of TMA), so what these tests establish is the *logic* of the
kernels, not their performance.  Nothing in the product imports this."""

import ctypes as ct
import hashlib
import os
import re
import subprocess
import tempfile

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_cache = os.path.join(tempfile.gettempdir(), 'pyfr_b200_cudaemu')

_sig_re = re.compile(
    r'extern "C" __global__ void\s*(?:__launch_bounds__\([^)]*\))?\s*'
    r'(\w+)\s*\(([^)]*)\)', re.S
)


def _wrapper(src):
    name, params = _sig_re.search(src).groups()
    params = [p.strip() for p in params.split(',') if p.strip()]
    types = [re.sub(r'\w+$', '', p).strip() for p in params]

    unpack = '\n'.join(f'    auto a{i} = *reinterpret_cast<{t} *>(args[{i}]);'
                       for i, t in enumerate(types))
    call = ', '.join(f'a{i}' for i in range(len(types)))
    threaded = int('__syncthreads' in src)

    return name, f'''
extern "C" void emu_entry(void **args, unsigned gx, unsigned gy, unsigned gz,
                          unsigned bx, unsigned by, unsigned bz)
{{
{unpack}
    emu_launch([&] {{ {name}({call}); }}, gx, gy, gz, bx, by, bz, {threaded});
}}
'''


def translate(src):
    from pyfr_b200.kernels.mul import _cpasync_src, _pipeline_src

    src = src.replace(_pipeline_src, '').replace(_cpasync_src, '')
    src = src.replace('asm volatile("fence.mbarrier_init.release.cluster;" '
                      '::: "memory");', '')
    src = re.sub(r'extern __shared__[^;]*;', '', src)
    src = re.sub(r'static __device__ __align__\((\d+)\) const (\w+)',
                 r'alignas(\1) static const \2', src)
    src = src.replace('#define UNROLL _Pragma("unroll")', '#define UNROLL')

    name, wrap = _wrapper(src)
    return name, f'#include "{_header()}"\n{src}\n{wrap}'


_flags = ['-std=c++17', '-O1', '-fPIC', '-pthread', '-w',
          '-fsanitize=alignment', '-fsanitize-undefined-trap-on-error']


def _header(tsan=None):
    """The execution-model header, copied next to a precompiled form of
    itself (parsing <thread>, <mutex>, ... is most of the compile time of a
    small kernel).  Keyed by content and build flavour."""
    if tsan is None:
        tsan = bool(os.environ.get('PYFR_B200_EMU_TSAN'))

    with open(os.path.join(_here, 'emu.h')) as f:
        text = f.read()

    key = hashlib.sha256(text.encode()).hexdigest()[:16]
    os.makedirs(_cache, exist_ok=True)
    hdr = os.path.join(_cache, f'emu-{key}{"-tsan" if tsan else ""}.h')

    if not os.path.exists(hdr + '.gch'):
        tmp = f'{hdr}.{os.getpid()}'
        with open(tmp, 'w') as f:
            f.write(text)
        os.replace(tmp, hdr)

        extra = ['-fsanitize=thread', '-g'] if tsan else []
        res = subprocess.run(['g++', *_flags, *extra, '-x', 'c++-header',
                              hdr, '-o', f'{tmp}.gch'],
                             capture_output=True, text=True)
        if res.returncode == 0:
            os.replace(f'{tmp}.gch', hdr + '.gch')

    return hdr


def compile_source(src):
    name, text = translate(src)
    key = hashlib.sha256(text.encode()).hexdigest()[:24]
    os.makedirs(_cache, exist_ok=True)
    so = os.path.join(_cache, f'{name}-{key}.so')

    # PYFR_B200_EMU_TSAN=1 builds the kernels with ThreadSanitizer (run the
    # tests with LD_PRELOAD=$(gcc -print-file-name=libtsan.so)): accesses
    # to shared/global memory by different CUDA threads that no barrier or
    # mbarrier orders are reported as data races
    tsan = bool(os.environ.get('PYFR_B200_EMU_TSAN'))
    if tsan:
        so = so[:-3] + '-tsan.so'

    if not os.path.exists(so):
        import threading
        uniq = f'{os.getpid()}-{threading.get_ident()}'
        cpp = f'{so[:-3]}.{uniq}.cpp'
        with open(cpp, 'w') as f:
            f.write(text)
        res = subprocess.run(
            ['g++', *_flags, '-shared',
             *(['-fsanitize=thread', '-g'] if tsan else []),
             '-o', f'{so}.{uniq}.tmp', cpp], capture_output=True, text=True
        )
        if res.returncode:
            raise RuntimeError(f'g++ failed for {name}:\n{res.stderr[:3000]}')
        os.replace(f'{so}.{uniq}.tmp', so)
        os.replace(cpp, so[:-3] + '.cpp')

    return so


# A system creates all of its kernels before it launches any: compile them
# concurrently and only wait for a shared object at its first launch
_pool = None


class _Module:
    def __init__(self, src):
        global _pool
        if _pool is None:
            from concurrent.futures import ThreadPoolExecutor
            _pool = ThreadPoolExecutor(max_workers=os.cpu_count() or 4)

        _header()                       # (precompiled once, not per worker)
        self._fut = _pool.submit(compile_source, src)
        self._entry = None

    @property
    def entry(self):
        if self._entry is None:
            self.lib = ct.CDLL(self._fut.result())
            self._entry = self.lib.emu_entry
            self._entry.restype = None
            self._entry.argtypes = [ct.c_void_p] + [ct.c_uint]*6

        return self._entry


class EmuRuntime:
    """The subset of pyfr_b200.lib.Runtime the backend uses."""

    dry = False
    emulated = True

    def __init__(self):
        self._bufs = {}
        self._mods = []
        self.nlaunch = 0
        self._cap = None            # operations of the graph being captured
        self.ncaptures = 0

    # -- stream capture -------------------------------------------------------
    # Modelled on CUDA graphs: while a capture is open nothing executes;
    # kernel parameters are copied *by value* into the captured node, so a
    # replay does not see later changes to the caller's argument storage
    # (the backend has to re-capture, which is what its dirty flags are for)
    def capture_begin(self, stream):
        if self._cap is not None:
            raise RuntimeError('nested stream capture')
        self._cap = []

    def end_capture(self, stream):
        ops, self._cap = self._cap, None
        self.ncaptures += 1
        return ops

    def graph_launch(self, graph, stream):
        if self._cap is not None:
            raise RuntimeError('graph launch during capture')
        for op in graph:
            op()

    def _do(self, op):
        if self._cap is not None:
            self._cap.append(op)
        else:
            op()

    # -- handles ------------------------------------------------------------
    def new_ptr(self, fn, *args):
        return fn(*args)

    def malloc(self, nbytes):
        buf = np.zeros(max(int(nbytes), 1) + 256, dtype=np.uint8)
        ptr = (buf.ctypes.data + 255)//256*256
        self._bufs[ptr] = buf
        return ptr

    malloc_host = malloc

    def free(self, ptr):
        self._bufs.pop(ptr, None)

    free_host = free

    def stream_create(self):
        return 1

    def stream_create_priority(self, high):
        return 2

    def event_create(self):
        return 3

    def _noop(self, *a):
        return None

    stream_destroy = stream_sync = device_sync = event_destroy = _noop
    event_record = event_sync = stream_wait_event = _noop
    module_unload = graph_destroy = function_set_dynamic_smem = _noop

    def elapsed_ms(self, a, b):
        return 0.0

    # -- data movement ------------------------------------------------------
    def memset(self, ptr, value, nbytes, stream):
        self._do(lambda: ct.memset(ptr, value, nbytes))

    def memcpy(self, dst, src, nbytes):
        if self._cap is not None:
            raise RuntimeError('synchronous copy during stream capture')
        ct.memmove(dst, src, nbytes)

    def memcpy_async(self, dst, src, nbytes, stream):
        self._do(lambda: ct.memmove(dst, src, nbytes))

    def memcpy2d_async(self, dst, dpitch, src, spitch, width, height, stream):
        def op():
            for r in range(height):
                ct.memmove(dst + r*dpitch, src + r*spitch, width)
        self._do(op)

    # -- kernels --------------------------------------------------------------
    def module_load(self, image):
        m = _Module(bytes(image).decode())
        self._mods.append(m)
        return m

    def module_get_function(self, module, name):
        return module

    def function_attrs(self, func):
        return dict(nregs=0, static_smem=0, local_bytes=0, max_threads=1024)

    def launch(self, func, gx, gy, gz, bx, by, bz, smem, stream, argv):
        if smem > 256*1024:
            raise RuntimeError('emulated shared memory exceeded')

        dims = tuple(map(int, (gx, gy, gz, bx, by, bz)))

        if self._cap is None:
            func.entry(ct.cast(argv, ct.c_void_p), *dims)
            self.nlaunch += 1
            return

        # Snapshot the parameters (every one is at most 8 bytes wide)
        n = len(argv)
        vals = (ct.c_uint64*n)()
        for i in range(n):
            ct.memmove(ct.addressof(vals) + 8*i, argv[i], 8)
        ptrs = (ct.c_void_p*n)(*[ct.addressof(vals) + 8*i for i in range(n)])

        def op(keep=(vals, ptrs)):
            func.entry(ct.cast(ptrs, ct.c_void_p), *dims)
            self.nlaunch += 1

        self._cap.append(op)

    def device_info(self):
        return dict(sm_count=148, cc=(10, 0), total_mem=0, free_mem=0,
                    smem_optin=232448)

    def __getattr__(self, name):
        def refuse(*a, **k):
            raise RuntimeError(f'b200_{name}: not available under emulation')
        return refuse


def install(monkeypatch):
    """Route B200Backend through the emulator."""
    import pyfr_b200.backend as bk
    import pyfr_b200.compiler as comp

    monkeypatch.setattr(bk, 'load_runtime',
                        lambda device=0, dry=False: EmuRuntime())
    monkeypatch.setattr(comp.KernelCompiler, 'cubin',
                        lambda self, src, name: src.encode())
