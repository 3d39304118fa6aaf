"""The C-ABI library builds, loads and exports every symbol include/bsk.h declares (no GPU needed)."""
import os
import re

from conftest import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "bsk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(bsk_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_listed():
    from bigseqkit_b200.api import ABI_SYMBOLS
    assert sorted(ABI_SYMBOLS) == _declared()


def test_library_exports_every_symbol():
    from bigseqkit_b200.api import LIB_PATH, Library
    assert os.path.exists(LIB_PATH), "run __graft_entry__.build() first"
    lib = Library()
    for s in _declared():
        assert hasattr(lib.cdll, s), s
    assert lib.cdll.bsk_version() >= 100


def test_create_reports_flag_errors_without_gpu():
    # option validation (the reference's Before()) happens before any CUDA call
    import ctypes as C
    from bigseqkit_b200.api import Library
    lib = Library()
    h = C.c_void_p()
    rc = lib.cdll.bsk_create(b"SeqTransform", b'{"LowerCase":true,"UpperCase":true}', -1, C.byref(h))
    assert rc == -1
    assert lib.cdll.bsk_create_error() == b"could not give both flags -l (--lower-case) and -u (--upper-case)"
    rc = lib.cdll.bsk_create(b"NoSuchOp", b"{}", -1, C.byref(h))
    assert rc == -1 and b"unknown operator" in lib.cdll.bsk_create_error()


def test_no_cpu_fallback_without_device():
    import ctypes as C
    from bigseqkit_b200.api import Library
    lib = Library()
    if lib.device_count() > 0:
        return
    h = C.c_void_p()
    rc = lib.cdll.bsk_create(b"SeqTransform", b"{}", -1, C.byref(h))
    assert rc == -3 and b"no CPU fallback" in lib.cdll.bsk_create_error()
