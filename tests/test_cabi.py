"""The C-ABI shared library: loads without a GPU, exports every symbol include/fgnn_b200.h
declares, the ctypes struct mirrors the C struct byte for byte, and argument validation works
(no compute calls here)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

import fgnn_b200
from fgnn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fgnn_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fgnn_[a-z0-9_]+)\s*\(", src)))


def test_library_built_and_loads():
    fgnn_b200.build()
    assert os.path.exists(_lib.LIB_PATH)
    lib = _lib.lib()
    assert lib.fgnn_version() == 200


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert "fgnn_mp_forward" in names and len(names) >= 10
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/fgnn_b200.h but not exported"
        assert n in _lib.EXPORTS, f"{n} has no ctypes prototype in _lib.EXPORTS"
    assert sorted(_lib.EXPORTS) == names


def test_struct_layout_matches_header(tmp_path):
    fields = [f[0] for f in _lib.MpArgs._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){',
            'printf("%zu\\n", sizeof(fgnn_mp_args));']
    prog += [f'printf("%zu\\n", offsetof(fgnn_mp_args, {f}));' for f in fields]
    prog += ['return 0;}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c11", "-o", str(exe), str(src)])
    out = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert out[0] == ctypes.sizeof(_lib.MpArgs)
    for f, off in zip(fields, out[1:]):
        assert getattr(_lib.MpArgs, f).offset == off, f


def test_halo_struct_layout_matches_header(tmp_path):
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){',
            'printf("%zu\\n", sizeof(fgnn_halo_args));', 'printf("%zu\\n", sizeof(fgnn_halo_job));']
    prog += [f'printf("%zu\\n", offsetof(fgnn_halo_args, {f[0]}));' for f in _lib.HaloArgs._fields_]
    prog += [f'printf("%zu\\n", offsetof(fgnn_halo_job, {f[0]}));' for f in _lib.HaloJob._fields_]
    prog += ['return 0;}']
    src = tmp_path / "layout_halo.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout_halo"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c11", "-o", str(exe), str(src)])
    out = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert out[0] == ctypes.sizeof(_lib.HaloArgs) and out[1] == ctypes.sizeof(_lib.HaloJob)
    offs = out[2:]
    for f, off in zip(_lib.HaloArgs._fields_, offs):
        assert getattr(_lib.HaloArgs, f[0]).offset == off, f[0]
    for f, off in zip(_lib.HaloJob._fields_, offs[len(_lib.HaloArgs._fields_):]):
        assert getattr(_lib.HaloJob, f[0]).offset == off, f[0]


def test_exchange_struct_layout_matches_header(tmp_path):
    fields = [f[0] for f in _lib.ExchangeArgs._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){',
            'printf("%zu\\n", sizeof(fgnn_exchange_args));']
    prog += [f'printf("%zu\\n", offsetof(fgnn_exchange_args, {f}));' for f in fields]
    prog += ['return 0;}']
    src = tmp_path / "layout_ex.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout_ex"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c11", "-o", str(exe), str(src)])
    out = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert out[0] == ctypes.sizeof(_lib.ExchangeArgs)
    for f, off in zip(fields, out[1:]):
        assert getattr(_lib.ExchangeArgs, f).offset == off, f


def test_exchange_validation_without_gpu():
    lib = _lib.lib()
    a = _lib.ExchangeArgs()
    assert lib.fgnn_exchange_forward(None, None) == _lib.ERR_INVALID_ARG
    a.world, a.rank, a.rows, a.J, a.O, a.row1 = 9, 0, 10, 1, 64, 10          # more ranks than the arena flags hold
    assert lib.fgnn_exchange_forward(ctypes.byref(a), None) == _lib.ERR_INVALID_ARG
    a.world = 2
    assert lib.fgnn_exchange_forward(ctypes.byref(a), None) == _lib.ERR_INVALID_ARG   # no counter / buffers


def test_enums_match_header():
    src = open(HEADER).read()
    for name, val in [("FGNN_AGG_MAX", _lib.AGG_MAX), ("FGNN_AGG_SOFTMAX", _lib.AGG_SOFTMAX),
                      ("FGNN_AGG_MEAN", _lib.AGG_MEAN), ("FGNN_AGG_NONE", _lib.AGG_NONE),
                      ("FGNN_ERR_INDEX_RANGE", _lib.ERR_INDEX_RANGE), ("FGNN_ERR_SHAPE", _lib.ERR_SHAPE),
                      ("FGNN_ERR_NO_DEVICE", _lib.ERR_NO_DEVICE), ("FGNN_KERNEL_TCGEN05", _lib.KERNEL_TCGEN05)]:
        m = re.search(rf"\b{name}\s*=\s*(-?\d+)", src)
        assert m and int(m.group(1)) == val, name


def _args(**kw):
    a = _lib.MpArgs()
    dummy = 0x1000
    a.x = a.idx = a.etype = a.filters = a.out = dummy
    a.B, a.N, a.M, a.K, a.C, a.O, a.T = 1, 4, 4, 2, 8, 8, 2
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def test_validation_without_gpu():
    lib = _lib.lib()
    assert lib.fgnn_mp_select_kernel(None) == _lib.ERR_INVALID_ARG
    assert lib.fgnn_mp_select_kernel(ctypes.byref(_args())) in (_lib.KERNEL_SIMT, _lib.KERNEL_TCGEN05)
    assert lib.fgnn_mp_select_kernel(ctypes.byref(_args(extension=3))) == _lib.ERR_INVALID_ARG
    assert lib.fgnn_mp_select_kernel(ctypes.byref(_args(K=0))) == _lib.ERR_INVALID_ARG
    assert lib.fgnn_mp_select_kernel(ctypes.byref(_args(extension=2, M=5))) == _lib.ERR_SHAPE
    assert lib.fgnn_mp_select_kernel(ctypes.byref(_args(x=None))) == _lib.ERR_INVALID_ARG
    assert lib.fgnn_mp_workspace_bytes(ctypes.byref(_args(kernel=_lib.KERNEL_SIMT))) == 0
    assert b"out of range" in lib.fgnn_strerror(_lib.ERR_INDEX_RANGE)
    assert isinstance(lib.fgnn_launch_count(), int)


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="only meaningful without a GPU")
def test_forward_without_device_fails_loudly():
    rc = _lib.lib().fgnn_mp_forward(ctypes.byref(_args()), None)
    assert rc in (_lib.ERR_NO_DEVICE, _lib.ERR_CUDA)
    with pytest.raises(_lib.FgnnError):
        _lib.check(rc, "mp_forward")
