import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EMU_SO = os.path.join(ROOT, "tools", "emu", "libbsk_emu.so")
CSRC = os.path.join(ROOT, "bigseqkit_b200", "csrc")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run with -m gpu on the B200 box)")


def _emu_stale():
    if not os.path.exists(EMU_SO):
        return True
    t = os.path.getmtime(EMU_SO)
    for dp, _, fs in os.walk(CSRC):
        if os.path.basename(dp) == "build":
            continue
        for f in fs:
            if f.endswith((".cu", ".h", ".cpp")) and os.path.getmtime(os.path.join(dp, f)) > t:
                return True
    return os.path.getmtime(os.path.join(ROOT, "include", "bsk.h")) > t


_libs = {}


def get_lib(kind):
    """kind == "gpu": the product library libbsk.so on a real GPU.
    kind == "emu": the development build of the SAME kernel sources on the host emulator
    (tools/emu, -DBSK_EMU) -- checks kernel logic in the GPU-less container; never shipped."""
    from bigseqkit_b200.api import Library
    if kind not in _libs:
        if kind == "emu":
            if _emu_stale():
                subprocess.check_call(["make", "-C", CSRC, "-s", "emu"])
            _libs[kind] = Library(EMU_SO)
        else:
            lib = Library()
            if lib.device_count() < 1:
                pytest.fail("libbsk.so loaded but no CUDA device is visible")
            _libs[kind] = lib
    return _libs[kind]


@pytest.fixture(params=[pytest.param("emu"), pytest.param("gpu", marks=pytest.mark.gpu)])
def lib(request):
    return get_lib(request.param)
