"""One warm-up + N timed run_contrack steps on a synthetic cube (for ncu launch lists / captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from contrack_b200 import Engine
T = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
a = torch.empty((T, bench.H, bench.W), dtype=torch.float32, device='cuda')
bench.synth_fill(a, 0, T)
f = torch.empty((T, bench.H, bench.W), dtype=torch.int32, device='cuda')
lat, lon = bench.grid(); w = bench.reference_weights(lat, lon)
e = Engine.get(0)
for k, v in [kv.split('=') for kv in sys.argv[3:]]:
    e.set_option(k, int(v))
for i in range(n + 1):
    e.run_contrack(a, w, 160, True, 0, 0.5, 5, True, out=f)
torch.cuda.synchronize()
print({k: round(v, 3) for k, v in e.stats().items()})
