"""CPU oracle for the nnx-ppo PPO hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``nnx_ppo_b200``) may
import this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do.

The oracle is a NumPy float32 *restatement* of the reference algorithm
(emiwar/nnx-ppo v0.2.1, files cited per function).  The reference is 100 % JAX /
flax.nnx / optax and none of those are installable in this image (no wheels, no
network), so it cannot be imported or executed here.

Pinning status
--------------
* pinned by the reference's own known-answer tests: ``gae`` (``ppo_test.py:229-264``),
  Normalizer moments (``normalizer_test.py:33-65``, ``factories_test.py:76-119``),
  replay consistency (``adapter_test.py:61-75``), DummyCounter-style bookkeeping;
* pinned by published third-party known answers: threefry2x32 (Random123 KATs) and the JAX
  ``split`` / ``fold_in`` values recorded in SURVEY.md App. B;
* **parity unpinned** for everything else (sampled actions, log-probs, losses, gradients,
  Adam, permutation indices): the reference ships no golden vectors for those, and it cannot be
  run here to generate any.  They are restated from the reference source plus the published
  semantics of jax / flax.nnx / optax, and cross-checked only internally (finite differences,
  torch.autograd in float64).
"""
