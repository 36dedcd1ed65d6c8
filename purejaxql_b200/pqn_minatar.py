"""PQN on MinAtar with the CNN Q-network — drop-in for purejaxql/pqn_minatar.py.

    python -m purejaxql_b200.pqn_minatar +alg=pqn_minatar alg.ENV_NAME=Breakout-MinAtar NUM_SEEDS=16

``make_train(config)`` keeps the reference's contract (pqn_minatar.py:89-431):
it mutates ``config`` (NUM_UPDATES, NUM_UPDATES_DECAY, TEST_NUM_STEPS), asserts
the minibatch divisibility, and returns ``train``.  The reference wraps
``train`` in ``jax.jit(jax.vmap(...))`` over ``rngs``; here ``train(rngs)`` takes
the ``[NUM_SEEDS, 2]`` key array directly and returns the same dict with a
leading seed axis: ``{"runner_state": (train_state, (obs, env_state),
test_metrics, rng), "metrics": {name: [S, NUM_UPDATES]}}``.
"""
from __future__ import annotations

from . import _runner, envs
from .engine import PQNEngine, prepare_config


def make_train(config):
    env, env_params = envs.make(config["ENV_NAME"])                  # :103-104
    prepare_config(config, env_params.max_steps_in_episode, allow_test_steps_override=False)   # :91-105
    engine = PQNEngine(config, network="cnn", flatten_obs=False)

    def train(rngs):
        return engine.train(rngs)

    train.engine = engine
    return train


def single_run(config):
    return _runner.single_run(config, make_train)


def tune(default_config):
    return _runner.tune(default_config, make_train)


def main(argv=None):
    return _runner.main(make_train, argv)


if __name__ == "__main__":
    main()
