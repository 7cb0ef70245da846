"""The class boundary with an `xarray` module present (tests/golden/xr_shim stands in for the wheel this image lacks):
the reference's own tests (tests/test_contrack.py:28-103) restated, in a subprocess because `contrack_b200.contrack`
decides at import time whether xarray exists."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(mode):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', '_xr_mode_worker.py'), mode], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'xarray-mode ok' in r.stdout


def test_container_with_xarray_module():
    _run('cpu')


@pytest.mark.gpu
def test_hot_path_with_xarray_module():
    _run('gpu')
