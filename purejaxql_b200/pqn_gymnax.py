"""PQN on gymnax classic control with the MLP Q-network — drop-in for
purejaxql/pqn_gymnax.py.

    python -m purejaxql_b200.pqn_gymnax +alg=pqn_cartpole NUM_SEEDS=8

Differences from pqn_minatar (as in the reference, pqn_gymnax.py:29-58,92-97):
MLP ``QNetwork(HIDDEN_SIZE, NUM_LAYERS)`` without the /255, the observation is
flattened (``FlattenObservationWrapper``), ``TEST_NUM_STEPS`` may be overridden
from the config, and there is REW_SCALE.
"""
from __future__ import annotations

from . import _runner, envs
from .engine import PQNEngine, prepare_config


def make_train(config):
    env, env_params = envs.make(config["ENV_NAME"], flatten_obs=True)      # :92-94
    prepare_config(config, env_params.max_steps_in_episode, allow_test_steps_override=True)    # :80-97
    engine = PQNEngine(config, network="mlp", flatten_obs=True)

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
