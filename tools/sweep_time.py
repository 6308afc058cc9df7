import os, sys, torch
sys.path.insert(0, '/root/repo')
from pyvb_b200 import PlateEngine
from tools.time_i8 import synth
N, D, q, miss = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
X = synth(N, D, q, miss, torch.device('cuda:0'))
e = PlateEngine(X, q, mode='B', keep_sigma=False); e.init_random(seed=5)
for _ in range(3): e.iterate_async()
torch.cuda.synchronize()
import time
ts=[]
for _ in range(5):
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); e.iterate_async(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print('sweep ms', min(ts), 'elbo', float(e.trace[(e.trace_pos-1)%e.trace.numel()].item()))
