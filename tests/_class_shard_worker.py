"""The class API on a time-sharded cube (tests/test_zz_examples.py): two contrack objects, each holding its own slice of the
time axis, call run_contrack(time_shard=...) collectively.  The two ranks are host threads with their own contexts and an
in-process communicator on one GPU."""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    from oracle import contrack_oracle as oracle
    from contrack import contrack
    from contrack_b200 import Dataset, Engine, sharded
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'anom_test.npz'))
    T = d['anom'].shape[0]
    cuts = [0, 6, T]
    ref = oracle.run_contrack(d['anom'], d['latitude'], d['longitude'], 150, '>=', 0.5, 5, True)
    comms = sharded.Comm.local_group(2)
    engines = [Engine(0), Engine(0)]
    objs, errs = [], [None, None]
    for r in range(2):
        lo, hi = cuts[r], cuts[r + 1]
        data = d['anom'][lo:hi] if r == 0 else torch.from_numpy(np.ascontiguousarray(d['anom'][lo:hi])).cuda()   # host / device
        ds = Dataset({'anom': (('time', 'latitude', 'longitude'), data)},
                     coords={'time': d['time'][lo:hi], 'latitude': d['latitude'], 'longitude': d['longitude']})
        c = contrack(ds=ds)
        c._engine = (lambda e: (lambda: e))(engines[r])           # (one process per GPU would use Engine.get(device))
        objs.append(c)

    def work(r):
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(torch.cuda.Stream(device=0)):
                objs[r].run_contrack('anom', 150, '>=', 0.5, 5, time_shard=(cuts[r], T), comm=comms[r])
                torch.cuda.current_stream().synchronize()
        except BaseException as e:                                  # noqa: BLE001
            errs[r] = e
            os._exit(4)                                             # (the other rank would wait for ever)
    th = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for r in range(2):
        got = objs[r]['flag'].values if hasattr(objs[r]['flag'], 'values') else np.asarray(objs[r]['flag'])
        if not np.array_equal(np.asarray(got), ref[cuts[r]:cuts[r + 1]]):
            print('rank %d MISMATCH' % r, flush=True)
            sys.exit(3)
    print('class sharded ok', flush=True)


if __name__ == '__main__':
    main()
