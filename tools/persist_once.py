"""One PPO-Lagrangian repeat through the persistent launch on a c2-shaped batch (for ncu captures):
python tools/persist_once.py [n_env=2048]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_ppo_scale_gpu import _collect  # noqa: E402

n_env = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
policy, batch, ob, actor, critics = _collect("SafetyCarCircle-v0", (256, 256), n_env, 0.3)
policy._target_kl = 1e9
for _ in range(2):
    np.random.seed(1)
    policy.learn(batch, batch_size=256, repeat=1)
torch.cuda.synchronize()
print("done", batch.n)
