"""examples/track_host.c: the C ABI from plain C99 (no CUDA header in the program).  Without a GPU the program must fail
loudly; on a B200 its flag cube must equal the oracle's on the very bytes it generated."""
import os
import subprocess

import numpy as np
import pytest

from oracle import contrack_oracle as oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, 'contrack_b200', 'lib')
T, H, W = 24, 91, 180


def build(tmp_path):
    exe = str(tmp_path / 'track_host')
    r = subprocess.run(['gcc', '-std=c99', '-O2', '-Wall', '-Wextra', '-pedantic', '-Werror', '-I', os.path.join(ROOT, 'include'),
                        os.path.join(ROOT, 'examples', 'track_host.c'), '-L', LIBDIR, '-lcontrack_b200', '-lm', '-o', exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def run(exe, tmp_path):
    a, f = str(tmp_path / 'anom.f32'), str(tmp_path / 'flag.i32')
    env = dict(os.environ, LD_LIBRARY_PATH=LIBDIR + os.pathsep + os.environ.get('LD_LIBRARY_PATH', ''))
    r = subprocess.run([exe, a, f], capture_output=True, text=True, env=env, timeout=600)
    return r, a, f


def expected(anom_path):
    x = np.fromfile(anom_path, np.float32).reshape(T, H, W)
    lat = (90.0 - 2.0 * np.arange(H)).astype(np.float32)
    lon = (2.0 * np.arange(W)).astype(np.float32)
    return oracle.run_contrack(x, lat, lon, 150, '>=', 0.5, 4, True)


def test_c_example_builds_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = build(tmp_path)
    r, a, _ = run(exe, tmp_path)
    ref = expected(a)                                   # (the cube is written before the device is touched)
    ids = np.unique(ref)
    # the story the example tells: the date-line crosser survives whole, the short-lived blob is removed, the jumper is cut
    # into two features
    assert len(ids) - 1 == 3 and ref[5, 60, 60] == 0 and ref[0, 25, 165 // 1 % W] != 0
    assert ref[3, 25, (165 + 6) % W] == ref[12, 25, (165 + 24) % W] != 0
    assert ref[3, 70, 103] != 0 and ref[20, 70, 170] != 0 and ref[3, 70, 103] != ref[20, 70, 170]
    if not torch.cuda.is_available():
        assert r.returncode == 2 and 'no usable CUDA device' in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_c_example_matches_the_oracle(tmp_path):
    exe = build(tmp_path)
    r, a, f = run(exe, tmp_path)
    assert r.returncode == 0, r.stderr
    ref = expected(a)
    got = np.fromfile(f, np.int32).reshape(T, H, W)
    assert np.array_equal(got, ref)
    assert r.stdout.startswith('features %d ' % (len(np.unique(ref)) - 1)), r.stdout


@pytest.mark.gpu
def test_class_api_on_a_time_sharded_cube():
    """contrack.run_contrack(time_shard=...) on two in-process ranks (one with host data, one with device data); in a
    subprocess with a time limit, because a rank that fails alone would leave the other one waiting."""
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', '_class_shard_worker.py')], capture_output=True, text=True,
                       cwd=ROOT, timeout=300)
    assert r.returncode == 0 and 'class sharded ok' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
