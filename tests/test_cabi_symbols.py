"""The C-ABI library loads and exports every symbol include/anml_b200.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

from anml_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "anml_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(anml_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_binding_covers_every_declared_symbol():
    assert sorted(_native.SIGNATURES) == declared_functions()


def test_version_and_error_string():
    lib = _native.load()
    assert lib.anml_version() == 200
    n = ctypes.c_int(-1)
    rc = lib.anml_device_count(ctypes.byref(n))
    assert rc in (_native.OK, _native.ECUDA)
    if rc != _native.OK:
        assert n.value == 0 and "cudaGetDeviceCount" in _native.last_error()


def test_argument_validation_without_a_device():
    lib = _native.load()
    h = ctypes.c_void_p()
    assert lib.anml_design_create(0, -1, 4, ctypes.byref(h)) == _native.EINVAL
    assert lib.anml_model_create(0, 10, 99, 1, ctypes.byref(h)) == _native.EINVAL
    assert "family" in _native.last_error()
    assert lib.anml_model_create(0, 10, 3, 1, ctypes.byref(h)) == _native.EINVAL  # meanvar needs 2 parameters


def test_no_cpu_fallback_when_no_device():
    import pytest

    if _native.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.Design(10, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.spline_design([0.5], [0.0, 1.0], 3)


def test_plain_c_program_links_and_runs(tmp_path):
    """gcc (C, not C++) compiles against the header and links the shared library."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        import pytest

        pytest.skip("gcc not available")
    exe = tmp_path / "link_check"
    lib_dir = os.path.dirname(_native.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cabi", "link_check.c"),
           "-L", lib_dir, "-lanml_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "version=200" in run.stdout
