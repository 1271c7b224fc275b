"""Drop-in check: run the REFERENCE's own unit tests, unchanged, against this package, with ``anml``
aliased to ``anml_b200`` (anml_b200/compat.py; ``xspline`` aliased to anml_b200.xspline).

The test files come from /root/reference/tests (the build container) or from oracle/_ref/tests, the
git-ignored copy that oracle/fetch_ref_tests.py makes so that the suite also travels to the GPU box.
Files that only need host logic replay everywhere; the ones that build a design matrix
(tests/parameter/test_main.py:94-153, tests/variable/test_spline.py:52-65, tests/getter/test_spline.py:18-28,
tests/getter/test_prior.py:80-83) need the device and replay under ``-m gpu``."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ["/root/reference/tests", os.path.join(ROOT, "oracle", "_ref", "tests")]
REF_TESTS = next((d for d in CANDIDATES if os.path.isdir(d)), None)
HOST_ONLY = ["data", "prior", "variable/test_main.py", "parameter/test_smoothmapping.py"]
GPU_FILES = ["parameter/test_main.py", "variable/test_spline.py", "getter"]

needs_ref = pytest.mark.skipif(REF_TESTS is None, reason="reference tests not present (run oracle/fetch_ref_tests.py where /root/reference exists)")


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


@needs_ref
def test_reference_host_tests_pass_against_anml_b200(tmp_path):
    res = _replay(tmp_path, HOST_ONLY)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert " passed" in res.stdout and "failed" not in res.stdout


@needs_ref
@pytest.mark.gpu
def test_reference_device_tests_pass_against_anml_b200(tmp_path):
    """Parameter.attach / get_params orders 0-2 / prior_dict shapes, SplineVariable.attach /
    get_design_mat, SplineGetter.get_spline, SplinePriorGetter.get_prior -- the reference's assertions on
    objects whose design matrices live on the GPU here."""
    res = _replay(tmp_path, GPU_FILES)
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert " passed" in res.stdout and "failed" not in res.stdout and "error" not in res.stdout.lower()
