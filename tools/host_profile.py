"""Where does the host spend its time in one c2 collect+update cycle?  cProfile over 3 cycles + CUDA-event split."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

cfg = bench.CONFIGS["c2"]
agent, trainer, col, buf, T = bench.build(cfg, "cuda:0", 0)
for _ in range(2):
    bench.one_cycle(trainer)
torch.cuda.synchronize()
t0 = time.time()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    bench.one_cycle(trainer)
torch.cuda.synchronize()
pr.disable()
print("wall per cycle %.1f ms" % ((time.time() - t0) / 3 * 1e3))
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
