#!/usr/bin/env python
"""The reference's actor loop (app/base.py:108-124) on the drop-in classes: two agents sharing one policy, a
replay recorder and per-episode tree resets.  Only the imports differ from the reference.

    python examples/dropin_episode.py --episodes 2 --sims 36
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from minitchess_alphazero_b200.agent import RoundRobinReferee, SimpleAlphaZeroAgent
from minitchess_alphazero_b200.environment import MinitChessEnvironment
from minitchess_alphazero_b200.erlyx_compat import BaseCallback, run_episodes
from minitchess_alphazero_b200.policy import Network, SimpleAlphaZeroPolicy


class InfoRecorder(BaseCallback):
    """exp/callbacks.py:31-54: the replay tuples the learner consumes."""

    def __init__(self, sink):
        self._sink = sink

    def on_episode_begin(self, initial_observation):
        self._rows, self._observation, self._reward = [], initial_observation, None

    def on_step_end(self, action, observation, reward, done):
        info = {'observation': self._observation, **action.info, 'action': int(action.action)}
        info['pi'] = info['pi'].tolist()
        self._rows.append(info)
        self._reward, self._observation = reward, observation

    def on_episode_end(self):
        reward = self._reward
        for info in reversed(self._rows):
            info['reward'] = reward
            reward = -reward
        self._sink.extend(self._rows)


class MonteCarloInit(BaseCallback):                         # exp/callbacks.py:57-62
    def __init__(self, agent):
        self._agent = agent

    def on_episode_begin(self, initial_observation):
        self._agent.init_mcts()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--episodes', type=int, default=1)
    ap.add_argument('--sims', type=int, default=36)          # app/base.py:25
    a = ap.parse_args()
    torch.manual_seed(0)
    np.random.seed(0)
    env = MinitChessEnvironment()
    policy = SimpleAlphaZeroPolicy(Network().eval())
    agents = [SimpleAlphaZeroAgent(environment=env, policy=policy, num_simulations=a.sims) for _ in range(2)]
    replay = []
    t0 = time.perf_counter()
    with torch.no_grad():
        run_episodes(env, RoundRobinReferee(agent_tuple=tuple(agents)), a.episodes,
                     callbacks=[InfoRecorder(replay), MonteCarloInit(agents[0]), MonteCarloInit(agents[1])], use_tqdm=False)
    dt = time.perf_counter() - t0
    print('%d episodes, %d plies, %.1f sims/s; last position %s reward %s' % (
        a.episodes, len(replay), len(replay) * a.sims / dt, replay[-1]['observation'], replay[-1]['reward']))


if __name__ == '__main__':
    main()
