"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the drone env / PPO hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it, and there only
as the checker (or as the thing timed as *the CPU baseline*), never as the shipped path.
The product package ``drone_rl_b200`` never imports from here.

Parity status
-------------
* env path (dynamics / obs / reward / done / reset / curriculum): PINNED -- the numpy
  restatement in ``drone_oracle.py`` is checked bit-for-bit against the reference's own
  ``drone.py`` / ``vectorized_drone.py`` (imported unmodified through ``ref_import.py``
  in the build container) and against the golden vectors in ``tests/golden/`` that were
  generated from the reference by ``tests/golden/make_golden.py``.
* vec-wrapper / VecMonitor / PPO path: PARITY UNPINNED -- the algorithm lives in
  stable-baselines3 (PyPI, un-vendored, un-pinned: reference environment.yaml:16), which
  is neither in /root/reference nor installed.  ``vecenv_oracle.py`` / ``ppo_oracle.py``
  restate its published algorithm; the reference holds no test or fixture for it.
"""
