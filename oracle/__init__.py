"""CPU oracle for the PQN hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker (or as the timed CPU
baseline), never as the thing shipped.  ``purejaxql_b200`` must not import it.

PARITY UNPINNED: the reference (mttga/purejaxql @47af6d7) has no tests, golden
vectors or fixtures, and its environment / PRNG arithmetic lives in third-party
packages that are absent from ``/root/reference`` and not installable here
(``gymnax==0.0.6``, ``jax>=0.4.16,<=0.4.38``, ``flax``, ``optax``; see
``pyproject.toml:27-31,51`` of the reference).  The modules here restate the
*published* algorithms of those pinned versions and are anchored on

* the Random123 Threefry-2x32-20 known-answer vectors (public KATs),
* publicly documented jax.random outputs for ``PRNGKey(0)`` (see
  ``tests/golden/README.md``),
* the reference's own call sites (file:line cited per function).

Modules: ``jax_prng`` (threefry / jax.random), ``gymnax_envs`` (MinAtar + classic control + wrappers), ``pqn_ref``
(Q-networks with the shipped layer_norm configuration, Q(lambda), RAdam, rollout, update step) and ``pqn_ref_norm``
(the NORM_TYPE=batch_norm / none and NORM_INPUT=True network variants -- oracle-first groundwork for SURVEY 8(f)
row 4; no kernel is built or claimed for them yet) and ``pqn_rnn_ref`` (the GRU network, in-loss Q(lambda) and
BPTT of ``pqn_rnn_gymnax.py`` -- same status).
"""
