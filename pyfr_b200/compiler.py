"""Kernel source -> sm_100a cubin, with an in-tree cache.

Counterpart of the reference's NVRTC wrapper with its on-disk cache
(``pyfr/backends/cuda/compiler.py:22-158``, ``pyfr/cache.py:57-63``).
Sources are generated at run time because operator constants are baked in;
the resulting cubins are cached under ``pyfr_b200/_kcache`` (in-tree, so a
cache populated by ``__graft_entry__.build()`` travels with the repository
snapshot).  On a cache miss the source is compiled with NVRTC through the
C-ABI when a device runtime is loaded, falling back to ``nvcc``; the
GPU-less build path always uses ``nvcc``.  Either way the target is
``sm_100a`` exactly -- no PTX JIT, no other architectures.
"""

import hashlib
import os
import shutil
import subprocess
import tempfile

ARCH = 'sm_100a'
_cache_dir = os.environ.get(
    'PYFR_B200_CACHE_DIR',
    os.path.join(os.path.dirname(os.path.abspath(__file__)), '_kcache')
)

_nvcc_flags = ['-gencode', 'arch=compute_100a,code=sm_100a', '-cubin', '-O3',
               '-lineinfo', '-std=c++17', '--fmad=true',
               '-Wno-deprecated-gpu-targets']
_nvrtc_flags = [f'--gpu-architecture={ARCH}', '-lineinfo', '--std=c++17',
                '--fmad=true', '-default-device']


def _nvcc():
    return shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'


def source_key(src):
    h = hashlib.sha256()
    h.update(ARCH.encode())
    h.update(' '.join(_nvcc_flags).encode())
    h.update(src.encode())
    return h.hexdigest()[:32]


def cache_path(src, name):
    return os.path.join(_cache_dir, f'{name}-{source_key(src)}.cubin')


def compile_nvcc(src, name, keep_src=False, extra=()):
    os.makedirs(_cache_dir, exist_ok=True)

    with tempfile.TemporaryDirectory() as td:
        cu, out = os.path.join(td, f'{name}.cu'), os.path.join(td, 'k.cubin')
        with open(cu, 'w') as f:
            f.write(src)

        cmd = [_nvcc(), *_nvcc_flags, *extra, '-o', out, cu]
        res = subprocess.run(cmd, capture_output=True, text=True)

        if res.returncode:
            dump = os.path.join(_cache_dir, f'failed-{name}.cu')
            shutil.copy(cu, dump)
            raise RuntimeError(f'nvcc failed for kernel {name!r} (source '
                               f'kept at {dump}):\n{res.stderr}')

        with open(out, 'rb') as f:
            return f.read(), res.stderr


class KernelCompiler:
    def __init__(self, rt):
        self.rt = rt
        self.stats = dict(hits=0, nvrtc=0, nvcc=0)

    def cubin(self, src, name):
        path = cache_path(src, name)

        try:
            with open(path, 'rb') as f:
                self.stats['hits'] += 1
                return f.read()
        except FileNotFoundError:
            pass

        image = None
        if not self.rt.dry and not os.environ.get('PYFR_B200_FORCE_NVCC'):
            try:
                image = self.rt.nvrtc(src, f'{name}.cu', _nvrtc_flags)
                self.stats['nvrtc'] += 1
            except Exception as e:
                if 'dlopen libnvrtc' not in str(e):
                    raise

        if image is None:
            image, _ = compile_nvcc(src, name)
            self.stats['nvcc'] += 1

        os.makedirs(_cache_dir, exist_ok=True)
        tmp = f'{path}.{os.getpid()}.tmp'
        with open(tmp, 'wb') as f:
            f.write(image)
        os.replace(tmp, path)

        if os.environ.get('PYFR_B200_KEEP_SRC'):
            with open(path.removesuffix('.cubin') + '.cu', 'w') as f:
                f.write(src)

        return image
