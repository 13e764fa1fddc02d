"""CPU tests of the C-ABI boundary: libpyfr_b200.so builds, loads and
exports exactly what include/pyfr_b200.h declares; without a device every
compute entry point fails loudly (there is no CPU fallback)."""

import ctypes as ct
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, 'include', 'pyfr_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(b200_\w+)\s*\(', hdr)))


def test_header_declares_entry_points():
    names = _declared()
    assert len(names) >= 40
    assert {'b200_init', 'b200_launch', 'b200_nccl_send',
            'b200_capture_begin', 'b200_last_error'} <= set(names)


def test_library_exports_every_declared_symbol(built):
    lib = ct.CDLL(os.path.join(ROOT, 'pyfr_b200', 'libpyfr_b200.so'))

    for name in _declared():
        assert hasattr(lib, name), f'{name} declared but not exported'


def test_binding_table_matches_header(built):
    from pyfr_b200.lib import exported_symbols

    assert sorted(exported_symbols) == _declared()


def test_no_device_is_an_error_not_a_fallback(built):
    """In a GPU-less process the runtime must refuse, with a message."""
    from pyfr_b200.lib import B200Error, Runtime

    rt = Runtime()
    has_gpu = os.path.exists('/dev/nvidia0')

    if has_gpu:
        pytest.skip('a CUDA device is present')

    with pytest.raises(B200Error) as ei:
        rt.init(0)
    assert str(ei.value)

    with pytest.raises(B200Error):
        rt.new_ptr(rt.malloc, 1024)


def test_dry_runtime_cannot_compute():
    """The build-time stand-in generates and compiles kernels but never
    executes anything: launches and copies raise."""
    from pyfr_b200.lib import B200NoDevice, DryRuntime

    rt = DryRuntime()
    for fn in ('launch', 'memcpy', 'memcpy_async', 'graph_launch',
               'nccl_send', 'memset'):
        with pytest.raises(B200NoDevice):
            getattr(rt, fn)()


def test_backend_requires_extension(monkeypatch, tmp_path):
    """A missing shared library is a hard error naming the build step."""
    from pyfr_b200 import lib

    with pytest.raises(lib.B200Error, match='build'):
        lib.Runtime(str(tmp_path / 'nonexistent.so'))
