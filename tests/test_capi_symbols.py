"""CPU: the shared libraries load and export every symbol the C headers declare (no compute calls)."""
import ctypes
import os
import re

import pytest

from tests.util import ROOT
from redtail_b200 import _lib


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def libs():
    paths = _lib.lib_paths()
    for p in paths.values():
        if not os.path.exists(p):
            import __graft_entry__
            __graft_entry__.build()
            break
    return {k: ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL) for k, p in paths.items()}


def test_kernel_header_symbols_exported(libs):
    names = _declared("redtail_b200.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(libs["kernels"], n), n
    assert set(names) == set(_lib.KERNEL_API), set(names) ^ set(_lib.KERNEL_API)


def test_engine_header_symbols_exported(libs):
    names = _declared("redtail_b200_engine.h")
    for n in names + ["createInferBuilder_INTERNAL", "createInferRuntime_INTERNAL"]:
        assert hasattr(libs["engine"], n), n
    assert set(names) | {"createInferBuilder_INTERNAL", "createInferRuntime_INTERNAL"} == set(_lib.ENGINE_API)


def test_bindings_load_and_report_version():
    assert b"sm_100a" in _lib.kernels_lib().rt_version()
    assert _lib.engine_lib().rt_stereo_last_error() is not None


def test_plugin_api_cpp_symbols_exported():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", _lib.lib_paths()["engine"]], capture_output=True, text=True).stdout
    for sym in ("redtail::tensorrt::IPluginContainer::create(nvinfer1::ILogger&)",
                "redtail::tensorrt::addElu(", "redtail::tensorrt::addCostVolume(", "redtail::tensorrt::addConv3D(",
                "redtail::tensorrt::addConv3DTranspose(", "redtail::tensorrt::addSlice(", "redtail::tensorrt::addTransform(",
                "redtail::tensorrt::addPad(", "redtail::tensorrt::addSoftargmax(",
                "redtail::tensorrt::StereoDnnPluginFactory::createPlugin(", "redtail::tensorrt::DimsUtils::getTensorSize("):
        assert sym in out, sym


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from redtail_b200 import StereoEngine
    from redtail_b200.ops import RedtailError
    with pytest.raises(RedtailError):
        StereoEngine("nvtiny", 161, 513, "/nonexistent")
    e = ctypes.c_void_p()
    rc = _lib.engine_lib().rt_stereo_create(b"nvtiny", 161, 513, 24, b"/nonexistent", 0, 1, ctypes.byref(e))
    assert rc != 0 and not e.value
