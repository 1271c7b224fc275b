"""Drop-in check: run the REFERENCE's own host-side unit tests against this
package, with ``anml`` aliased to ``anml_b200`` (anml_b200/compat.py).  Only
possible where /root/reference exists (the build container); skipped elsewhere.
Test files that need a design matrix (GPU) are replayed by test_gpu_api.py."""
import os
import subprocess
import sys
import textwrap

import pytest

REF_TESTS = "/root/reference/tests"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_ONLY = ["data", "prior", "variable/test_main.py", "parameter/test_smoothmapping.py"]
GPU_FILES = ["parameter/test_main.py", "variable/test_spline.py", "getter"]


def _replay(tmp_path, targets):
    import shutil

    dst = tmp_path / "tests"
    shutil.copytree(REF_TESTS, dst)
    (tmp_path / "conftest.py").write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        from anml_b200.compat import install_as_anml
        install_as_anml()
    """))
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", str(tmp_path)]
    cmd += [str(dst / t) for t in targets]
    return subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True)


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference tests not present")
def test_reference_host_tests_pass_against_anml_b200(tmp_path):
    res = _replay(tmp_path, HOST_ONLY)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert " passed" in res.stdout and "failed" not in res.stdout
