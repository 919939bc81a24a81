"""Host enqueue time vs GPU time of one PPO repeat (is the C launch loop the bottleneck?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fsrl_b200 import _lib

agent, trainer, col, buf, T = bench.build("cuda:0", 0)
bench.one_cycle(trainer)
torch.cuda.synchronize()
orig = _lib.lib.fsrl_ppo_lag_epoch
rec = []
class Wrap:
    def __call__(self, *a):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = orig(*a)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        rec.append((t1 - t0, t2 - t0))
        return r
_lib.lib.fsrl_ppo_lag_epoch = Wrap()
bench.one_cycle(trainer)
for enq, tot in rec:
    print(f"enqueue {enq*1e3:.1f} ms   enqueue+gpu {tot*1e3:.1f} ms")
