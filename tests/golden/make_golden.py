#!/usr/bin/env python3
"""Regenerates tests/golden/*.  Run in the BUILD container only (needs /root/reference, which is absent on the GPU box).

1. anom_test.npz   -- the reference's own test fixture (tests/test_data/anom_test.nc, 11x181x360 float32 'anom').
   No HDF5 reader is installed; the variable is stored contiguous/uncompressed as the last 11*181*360*4 bytes of
   the file (sha256 of file and of the extracted array are asserted below).
2. golden.json     -- expected results of the reference algorithm (restated oracle, oracle/contrack_oracle.py) on the
   fixture and on seeded synthetic cubes: sha256 of the flag array as C-order '<i4', the ids, non-zero cell count.
   The four fixture rows are the values SURVEY.md section 8(c) recorded independently.
3. flags_*.npz     -- full expected flag arrays (compressed) for the fixture cases.
4. lifecycle_fixture.json -- the run_lifecycle table of the reference test's run on the fixture.
"""
import hashlib, json, os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import contrack_oracle as oracle      # noqa: E402
from _synth import synth_cube, regular_grid       # noqa: E402

REF_NC = '/root/reference/tests/test_data/anom_test.nc'


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).astype('<i4').tobytes()).hexdigest()


def main():
    raw = open(REF_NC, 'rb').read()
    assert hashlib.sha256(raw).hexdigest().startswith('97acd756'), 'unexpected fixture file'
    a = np.frombuffer(raw[-11 * 181 * 360 * 4:], dtype='<f4').reshape(11, 181, 360).copy()
    assert hashlib.sha256(a.tobytes()).hexdigest().startswith('8a8d1c01')
    lat = np.arange(90, -91, -1, dtype=np.float32)
    lon = np.arange(360, dtype=np.float32)
    # the file's time axis: 11 daily steps, "days since 2016-10-02" 
    time = np.datetime64('2016-10-02') + np.arange(11).astype('timedelta64[D]')
    np.savez_compressed(os.path.join(HERE, 'anom_test.npz'), anom=a, time=time.astype('datetime64[ns]'), latitude=lat,
                        longitude=lon)

    gold = {'fixture': [], 'synthetic': [], 'quirk': []}
    flags = {}
    for thr, ov, pers, two in [(150, .5, 5, False), (150, .5, 5, True), (160, .5, 5, True), (100, .7, 3, True)]:
        f = oracle.run_contrack(a, lat, lon, thr, '>=', ov, pers, two)
        key = 'thr%d_ov%02d_p%d_%s' % (thr, int(ov * 10), pers, 'two' if two else 'one')
        flags[key] = f.astype(np.int32)
        gold['fixture'].append(dict(key=key, threshold=thr, gorl='>=', overlap=ov, persistence=pers, twosided=two,
                                    ids=[int(i) for i in np.unique(f)[1:]], nonzero=int((f > 0).sum()),
                                    t_id_pairs=int(sum(len(np.unique(f[t])) - 1 for t in range(f.shape[0]))),
                                    sha256=sha(f)))
    np.savez_compressed(os.path.join(HERE, 'flags_fixture.npz'), **flags)

    # run_lifecycle (contrack.py:799-907) on the reference test's own run (tests/test_contrack.py:93-103: 28 rows, 3 flags)
    rows = oracle.run_lifecycle(flags['thr150_ov05_p5_one'], a, lat, lon, time)
    assert len(rows) == 28 and len({r[0] for r in rows}) == 3
    json.dump({'source': 'oracle.run_lifecycle on the reference fixture, run_contrack(150, >=, 0.5, 5, twosided=False); '
                         'Intensity/Size as float.hex()',
               'rows': [[int(r[0]), r[1], int(r[2]), int(r[3]), float(r[4]).hex(), float(r[5]).hex()] for r in rows]},
              open(os.path.join(HERE, 'lifecycle_fixture.json'), 'w'), indent=0)

    # BASELINE.json configs[0]: synthetic 30x91x180, threshold 160, overlap 0.5, persistence 5, twosided
    for seed, T, H, W, sig, thr, gorl, ov, pers, two in [
            (1, 30, 91, 180, (2.5, 3, 5), 160, '>=', .5, 5, True),
            (1, 30, 91, 180, (2.5, 3, 5), 160, '>=', .5, 5, False),
            (1, 30, 91, 180, (2.5, 3, 5), -160, '<', .5, 5, True),
            (3, 40, 91, 180, (1.5, 2, 3), 120, '>', .3, 3, True),
            (5, 24, 181, 360, (2.0, 4, 6), 150, 'ge', .7, 4, True)]:
        x = synth_cube(seed, T, H, W, sig)
        la, lo = regular_grid(H, W)
        f = oracle.run_contrack(x, la, lo, thr, gorl, ov, pers, two)
        gold['synthetic'].append(dict(seed=seed, shape=[T, H, W], sigma=list(sig), threshold=thr, gorl=gorl,
                                      overlap=ov, persistence=pers, twosided=two,
                                      input_sha256=hashlib.sha256(x.tobytes()).hexdigest(),
                                      ids=[int(i) for i in np.unique(f)[1:]], nonzero=int((f > 0).sum()),
                                      sha256=sha(f)))

    # stale-bounding-box seam quirk (contrack.py:753-763): seeds where the reference differs from a clean periodic union
    for seed in [1207, 1219, 1280, 1288, 1339, 1367, 1396]:
        x = synth_cube(seed, 12, 24, 16, (1.5, 2, 2))
        f = oracle.track_persistence((x >= 60).astype(int), 1)
        gold['quirk'].append(dict(seed=seed, shape=[12, 24, 16], sigma=[1.5, 2, 2], threshold=60, persistence=1,
                                  ids=[int(i) for i in np.unique(f)[1:]], sha256=sha(f)))
    json.dump(gold, open(os.path.join(HERE, 'golden.json'), 'w'), indent=1)
    print('wrote', os.listdir(HERE))


if __name__ == '__main__':
    main()
