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
* The update paths -- ``PPOLagrangian / CPO / TRPOLagrangian / FOCOPS / SACLagrangian /
  DDPGLagrangian.learn`` -- are pinned against the reference's own classes as well:
  ``oracle/make_golden_policies.py`` (this container only) registers thin ``tianshou`` /
  ``gymnasium`` shims (attribute containers, spaces, plain ``torch.nn`` modules), imports the REAL
  ``fsrl.policy.*`` from ``/root/reference`` and drives ``learn()`` on CPU on seeded batches;
  the logged per-step statistics and final weights are the fixtures
  ``tests/golden/policy_*_golden.npz`` that ``oracle/{ppo,cpo,trpo,focops,offpolicy}.py`` must
  reproduce (CPO: dual cases 0-3; PPO: dual clip + value clip; SAC: auto / fixed alpha with the
  reference's own reparameterisation noise recorded).
* The same generator pins host-side pieces of the drop-in surface against the reference's own code:
  the ``compute_gae_returns`` / ``compute_nstep_returns`` glue (value mask, end flags, reward
  normalisation), the trainers + ``BaseLogger`` (call traces, returned statistics, file bytes;
  scenario in ``oracle/trainer_scenario.py``), every config dataclass, run naming, and
  ``map_action`` / ``map_action_inverse``, and ``FastCollector.collect`` itself (episode counting,
  surplus-env rule, reset order: nine scenarios over the numpy env twin, incl. scripted terminations).
* Still **unpinned** by executable reference code (``tianshou`` 0.5.0 itself is absent, no
  network): ``Batch.split`` ordering and ``VectorReplayBuffer`` ring-index semantics -- restated from SURVEY.md 2.3 /
  Appendix C and anchored on the reference's call sites.
* The environment dynamics (pybullet / mujoco) cannot be reproduced at all; the device
  env is *our* documented model and ``oracle/envs.py`` is its CPU twin.
"""
