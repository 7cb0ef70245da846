"""Worker of tests/test_sharded_cpu.py: one rank of a world_size-N gloo job (CPU only).

Each rank builds the tables of ITS time shard (+ the halo plane) with numpy (tests/_tables_np.py), then runs exactly the
host-side code of the multi-GPU path -- pack / all-gather / merge_views / ct_host_tables_fast with the collective plane
fetch -- and checks its planes of the result against the oracle on the whole cube."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch.distributed as dist
    from oracle import contrack_oracle as oracle
    from contrack_b200 import sharded
    import _shard_np as shard_np
    from _common import row_weights
    from _synth import synth_cube, regular_grid
    from _tables_np import build_tables, legacy_to_view
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    cases = [(1396, 12, 24, 16, (1.5, 2, 2), 60, 0.0, 1, False),       # stale-box split: needs the collective fetch
             (1003, 12, 24, 16, (1.5, 2, 2), 60, 0.5, 2, True),
             (7, 17, 30, 40, (1.5, 3, 4), 80, 0.4, 3, True),
             (9, 9, 19, 33, (1.0, 2, 3), 50, 0.6, 2, False)]
    fetches = 0
    for seed, T, H, W, sig, thr, ov, pers, two in cases:
        x = synth_cube(seed, T, H, W, sig)
        lat = (80 - np.arange(H) * 2.0).astype(np.float32)
        lon = (np.arange(W) * 2.0).astype(np.float32)
        if seed == 7:
            lat = np.linspace(90, -90, H).astype(np.float32)            # pole rows: special-row bookkeeping across ranks
            lat = (90 - np.arange(H) * (180.0 / (H - 1))).astype(np.float32)
        w = row_weights(lat, lon) if seed != 7 else oracle.weight_grid(
            lat, oracle.resolution(lat, True), oracle.resolution(lon, True), W)[:, 0].copy()
        ref = oracle.run_contrack(x, lat, lon, thr, '>=', ov, pers, two, force=True)
        t0, t1 = sharded.shard_bounds(T, world)[rank]
        has_prev = 1 if rank > 0 else 0
        mask = x[t0 - has_prev:t1] >= thr
        tb = build_tables(mask, w)
        mine = legacy_to_view(tb, has_prev, t0)
        views = [shard_np.unpack_view(b) for b in shard_np.allgather_bytes(shard_np.pack_view(mine))]
        g, offs = shard_np.merge_views(views)
        bounds = sharded.shard_bounds(T, world)

        def fetch(t):
            nonlocal fetches
            fetches += 1
            owner = next(r for r, (a, b) in enumerate(bounds) if a <= t < b)
            obj = [None]
            if rank == owner:
                lp = t - t0 + has_prev
                a, b = tb['plane_run_ptr'][lp], tb['plane_run_ptr'][lp + 1]
                obj = [(tb['run_y'][a:b], tb['run_x0'][a:b], tb['run_x1'][a:b],
                        (tb['run_comp'][a:b].astype(np.int64) + offs[rank]).astype(np.uint32))]
            dist.broadcast_object_list(obj, src=owner)
            return obj[0]

        val, overrides, stats = shard_np.host_tables_fast(T, H, W, w, g, ov, pers, two, fetch=fetch)
        # paint the own planes from the label images of build_tables
        flag = np.zeros((t1 - t0, H, W), np.int32)
        for lp in range(has_prev, tb['T']):
            lab, o = tb['labs'][lp], tb['offs'][lp]
            n = tb['offs'][lp + 1] - o
            lut = np.concatenate([[0], val[o + offs[rank]:o + offs[rank] + n]]).astype(np.int32)
            flag[lp - has_prev] = lut[lab]
        for (t, y, a, b, v) in overrides:
            if t0 <= t < t1:
                flag[t - t0, y, a:b] = v
        if not np.array_equal(flag, ref[t0:t1]):
            print('rank %d: MISMATCH in case seed=%d' % (rank, seed), flush=True)
            sys.exit(3)
        assert stats[0] == len(np.unique(ref)) - 1
    if fetches == 0:
        print('rank %d: the collective plane fetch was never exercised' % rank, flush=True)
        sys.exit(4)
    dist.barrier()
    dist.destroy_process_group()
    print('rank %d ok (%d plane fetches)' % (rank, fetches), flush=True)


if __name__ == '__main__':
    main()
