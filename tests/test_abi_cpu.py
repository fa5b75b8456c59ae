"""The C-ABI shared library loads without a GPU and exports every symbol include/edgegan_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "edgegan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from edgegan_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    from edgegan_b200 import _lib
    assert set(_lib.exported_symbols()) == set(declared_symbols())
    lib = _lib.load()
    assert lib.eg_abi_version() == 1


def test_argument_errors_are_reported_not_thrown():
    from edgegan_b200 import _lib
    lib = _lib.load()
    rc = lib.eg_fill(None, 10, 1.0, None)          # NULL destination
    assert rc < 0
    assert b"invalid argument" in lib.eg_last_error()
    s = _lib.ConvShape(1, 4, 4, 3, 2, 2, 8, 4, 4, 2, 5, 1)      # pad_t >= KH
    assert lib.eg_conv2d_fwd(ctypes.byref(s), None, None, None, None, 0, None) < 0


def test_product_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from edgegan_b200.ops import DeviceOps
    with pytest.raises(RuntimeError):
        DeviceOps()


def test_library_has_no_unresolved_internal_symbols():
    """internal cross-file calls are plain C++ declarations repeated per .cu file: a signature that drifts in one file
    links as an undefined symbol of a shared library and only fails when first called -- bind everything now"""
    import ctypes
    import os
    from edgegan_b200 import _lib
    ctypes.CDLL(_lib.LIB_PATH, mode=os.RTLD_NOW | os.RTLD_LOCAL)


def test_descriptor_struct_layouts_match_the_header():
    """ctypes mirrors of the descriptor structs (include/edgegan_b200.h: eg_filter_desc, eg_sn_desc, eg_conv_shape) have the
    C layout: field order, offsets and total size"""
    import ctypes as C
    from edgegan_b200 import _lib
    assert C.sizeof(_lib.FilterDesc) == 24 and _lib.FilterDesc.taps.offset == 8 and _lib.FilterDesc.Co.offset == 16
    names = ["W", "u", "Wbar", "ws", "G", "gW", "Wa", "Wi", "Ga", "Gi"]
    assert [f[0] for f in _lib.SnDesc._fields_] == names + ["K", "C", "cin", "hd"]
    assert C.sizeof(_lib.SnDesc) == 10 * 8 + 4 * 4
    for i, n in enumerate(names):
        assert getattr(_lib.SnDesc, n).offset == 8 * i
    assert _lib.SnDesc.K.offset == 80 and _lib.SnDesc.hd.offset == 92
    hdr = open(os.path.join(ROOT, "include", "edgegan_b200.h")).read()
    body = hdr[hdr.index("typedef struct {\n    const float* W;"):hdr.index("} eg_sn_desc;")]
    order = re.findall(r"\b(W|u|Wbar|ws|G|gW|Wa|Wi|Ga|Gi|K|C|cin|hd)\b[;,]", body)
    assert order == names + ["K", "C", "cin", "hd"], order
