"""One rank of the NCCL test of the time-sharded path (tests/test_gpu_sharded.py::test_sharded_torchrun_two_gpus):
ct_run_contrack_sharded over an NCCL communicator made inside the library (unique id distributed through torch.distributed)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    import torch.distributed as dist
    from oracle import contrack_oracle as oracle
    from contrack_b200 import Engine, sharded
    from _common import row_weights
    from _synth import synth_cube, regular_grid
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    eng = Engine.get(local)
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'anom_test.npz'))
    cases = [(d['anom'], d['latitude'], d['longitude'], 150, 0.5, 5, True),
             (synth_cube(1396, 12, 24, 16, (1.5, 2, 2)),) + regular_grid(24, 16) + (60, 0.0, 1, False),
             (synth_cube(5, 24, 181, 360, (2.0, 4, 6)),) + regular_grid(181, 360) + (150, 0.7, 4, True)]
    seen = set()
    for p2p in (1, 0, 1):                      # peer windows (NVLink stores + flag words), ncclAllGather, peer windows again
        eng.set_option('p2p', p2p)
        for x, lat, lon, thr, ov, pers, two in cases:
            w = row_weights(lat, lon)
            ref = oracle.run_contrack(x, lat, lon, thr, '>=', ov, pers, two)
            t0, t1 = sharded.shard_bounds(x.shape[0], world)[rank]
            xl = torch.from_numpy(np.ascontiguousarray(x[t0:t1])).cuda()
            flag, n, stats = sharded.run_contrack_sharded(eng, xl, t0, x.shape[0], w, thr, True, 0, ov, pers, two)
            torch.cuda.synchronize()
            seen.add((p2p, stats.get('p2p')))
            if not np.array_equal(flag.cpu().numpy(), ref[t0:t1]) or n != len(np.unique(ref)) - 1:
                print('rank %d MISMATCH (p2p=%d)' % (rank, p2p), flush=True)
                sys.exit(3)
    print('rank %d transports (asked, used): %s' % (rank, sorted(seen)), flush=True)
    # a rank that enters the call seconds after the others (host work, I/O) is waited for, as a collective would
    import time
    if rank == world - 1:
        time.sleep(6.0)
    x, lat, lon, thr, ov, pers, two = cases[0]
    t0, t1 = sharded.shard_bounds(x.shape[0], world)[rank]
    flag, n, _ = sharded.run_contrack_sharded(eng, torch.from_numpy(np.ascontiguousarray(x[t0:t1])).cuda(), t0, x.shape[0],
                                              row_weights(lat, lon), thr, True, 0, ov, pers, two)
    torch.cuda.synchronize()
    if not np.array_equal(flag.cpu().numpy(), oracle.run_contrack(x, lat, lon, thr, '>=', ov, pers, two)[t0:t1]):
        print('rank %d MISMATCH (late rank)' % rank, flush=True)
        sys.exit(3)
    dist.barrier()
    dist.destroy_process_group()
    print('rank %d ok' % rank, flush=True)


if __name__ == '__main__':
    main()
