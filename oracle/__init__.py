"""CPU oracle for the FSRL hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (numpy / torch-CPU / plain C), the algorithm of the
reference's hot path (liuzuxin/FSRL).  It exists to *check* the
CUDA product in ``fsrl_b200``; nothing under ``fsrl_b200/`` may import, link or execute
it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it.

Parity pinning
--------------
* ``gae_return`` / ``nstep_return`` / ``LagrangianOptimizer`` are pinned against the
  reference's *own* code: ``oracle/make_golden.py`` AST-extracts / imports them from
  ``/root/reference`` (read-only, this container only) and writes the fixtures under
  ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.
* Everything that needs ``tianshou`` / ``gymnasium`` / ``pybullet`` (absent, no network)
  is a restatement that follows the cited reference lines; for those pieces parity is
  **unpinned** by executable reference code (the reference has no numeric tests,
  SURVEY.md F7) and is anchored on the reference's call sites only.
* The environment dynamics (pybullet / mujoco) cannot be reproduced at all; the device
  env is *our* documented model and ``oracle/envs.py`` is its CPU twin.
"""
