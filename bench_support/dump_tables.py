import sys, os
sys.path.insert(0,'/root/repo')
import bench, torch, numpy as np
from contrack_b200 import Engine
T=int(sys.argv[1])
a=torch.empty((T,bench.H,bench.W),dtype=torch.float32,device='cuda')
bench.synth_fill(a,0,T)
lat,lon=bench.grid(); w=bench.reference_weights(lat,lon)
os.environ['CT_DUMP_TABLES']='gpurun_out/tables_%d.bin'%T
os.makedirs('gpurun_out',exist_ok=True)
e=Engine.get(0)
f,n=e.run_contrack(a,w,160,True,0,0.5,5,True)
print(n,e.stats())
