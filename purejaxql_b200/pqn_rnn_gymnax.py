"""Recurrent (GRU) PQN on gymnax classic control — drop-in for purejaxql/pqn_rnn_gymnax.py.

    python -m purejaxql_b200.pqn_rnn_gymnax +alg=pqn_rnn_cartpole NUM_SEEDS=4

``make_train(config)`` keeps the reference's contract (pqn_rnn_gymnax.py:117-560): config mutation (NUM_UPDATES,
NUM_UPDATES_DECAY, TEST_NUM_STEPS), ``RNNQNetwork`` (MLP trunk -> one-hot last action -> scanned GRU with done-resets ->
Q head), a memory of MEMORY_WINDOW + NUM_STEPS transitions warmed up with random actions, minibatches over ENVS (whole
trajectories) and the Q(lambda) targets computed inside the loss from the window's own q values.  As in the other
scripts ``train(rngs)`` takes the ``[NUM_SEEDS, 2]`` key array natively.
"""
from __future__ import annotations

from . import _runner, envs
from .engine import prepare_config
from .engine_rnn import PQNRnnEngine


def make_train(config):
    if config["ENV_NAME"] == "MemoryChain-bsuite":
        raise NotImplementedError("MemoryChain-bsuite is not built (CartPole-v1 / Acrobot-v1 are)")
    env, env_params = envs.make(config["ENV_NAME"], flatten_obs=True)      # :134-139
    prepare_config(config, env_params.max_steps_in_episode, allow_test_steps_override=True)    # :119-132,140
    engine = PQNRnnEngine(config)

    def train(rngs):
        return engine.train(rngs)

    train.engine = engine
    return train


def single_run(config):
    return _runner.single_run(config, make_train, alg_file_name="pqn_rnn")


def main(argv=None):
    return _runner.main(make_train, argv)


if __name__ == "__main__":
    main()
