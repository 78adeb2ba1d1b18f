"""CPU tier: the C-ABI shared library loads and exports every symbol include/vbhem_b200.h
declares; without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from vbhem_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vbhem_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vbh[em]m_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built():
    assert os.path.exists(_lib.SO_PATH), "run __graft_entry__.build()"


def test_header_symbols_exported_and_bound():
    lib = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(syms)


def test_stats_layout_helpers_need_no_gpu():
    lib = _lib.lib()
    K, r, d = 64, 8, 2
    n = lib.vbhem_stats_size(K, r, d, 1, 1)
    assert n == K * (1 + 8 + 64 + 8 + 16 + 32) + 2      # 64 x 129 + 2 doubles
    off = (C.c_size_t * 7)()
    assert lib.vbhem_stats_offsets(K, r, d, 1, 1, off) == 0
    assert list(off) == [0, K, K + 8 * K, K + 72 * K, K + 80 * K, K + 96 * K, K + 128 * K]
    assert lib.vbhem_version() >= 100


def _no_gpu():
    n = C.c_int()
    return _lib.lib().vbhem_device_count(C.byref(n)) != 0 or n.value == 0


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    from vbhem_b200.api import BaseSet
    from vbhem_b200 import synth
    base, _ = synth.make_base_set(4, S=2, d=2, C=2, seed=1)
    with pytest.raises(_lib.VbhemError) as ei:
        BaseSet(base)
    assert ei.value.code == -2 and "no CUDA device" in ei.value.msg
    from vbhem_b200.api import MultiBaseSet
    with pytest.raises(_lib.VbhemError) as ei:
        MultiBaseSet(base, devices=[0, 1])
    assert ei.value.code == -2 and "no CUDA device" in ei.value.msg


def test_nccl_is_loaded_at_run_time_not_linked():
    """The multi-GPU entry points resolve NCCL with dlopen: the library itself must not depend on libnccl, so that
    a single-GPU MATLAB host can load it without NCCL installed."""
    import subprocess
    out = subprocess.run(["ldd", _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "nccl" not in out
    uid = (C.c_ubyte * 128)()
    rc = _lib.lib().vbhem_multi_unique_id(uid)
    assert rc in (0, -2)                                     # -2: no libnccl.so.2 on this box
    if rc == 0:
        assert any(bytes(uid))


def test_argument_checks_precede_device_use():
    """Malformed inputs are rejected with the reference's error id before anything touches CUDA."""
    lib = _lib.lib()
    h = C.c_void_p()
    z = np.zeros(4)
    p = z.ctypes.data_as(_lib.dp)
    sb = np.array([3], np.int32)
    rc = lib.vbhem_pack_base(0, 1, sb.ctypes.data_as(_lib.ip), 2, 2, 1, p, p, p, p, None, C.byref(h))
    assert rc == -1 and b"vbhem_mex:invalidinput" in lib.vbhem_last_error()
    rc = lib.vbhem_pack_base(0, 1, sb.ctypes.data_as(_lib.ip), 4, 2, 7, p, p, p, p, None, C.byref(h))
    assert rc == -1 and b"covmode" in lib.vbhem_last_error()
    rc = lib.vbhem_estep(None, 1, sb.ctypes.data_as(_lib.ip), 4, None, 1, 5, p, p, p, p, p, p, None, None, p)
    assert rc == -1 and b"handle is NULL" in lib.vbhem_last_error()
