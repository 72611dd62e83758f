"""CPU oracle for the DQN train-step hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker or as the timed
CPU baseline.  ``deepqnetwork_b200`` never imports this package: the product
path fails loudly when the CUDA library is missing.

Pinning status
--------------
* replay / PER half (``sumtree_ref``, ``replay_ref``, ``sumtree_ref.c``):
  **pinned** against executions of the reference's own ``rl.data.*`` modules
  (``oracle/make_golden.py`` imports them unmodified from ``/root/reference``
  and writes ``tests/golden/*.npz``; ``tests/test_oracle_*.py`` replay them).
* NN half (``nn_ref``): **parity unpinned**.  TensorFlow is an un-vendored pip
  dependency (``Pipfile:13`` tensorflow==1.12.3, ``Pipfile.lock:256`` 1.9.0)
  that is not installed and not installable here, and the reference has no
  tests or golden vectors.  ``nn_ref`` restates the published TF semantics the
  reference's call sites rely on (SAME padding, NHWC flatten, huber_loss,
  ApplyAdam, ApplyMomentum) in torch-CPU fp32/fp64.
"""
