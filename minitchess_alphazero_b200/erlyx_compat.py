"""The erlyx names the reference's self-play path imports (exp/agent.py:1-2, exp/environment.py:2,4,
exp/policy.py:1).  Uses the real `erlyx` when it is installed, otherwise minimal equivalents, so
the facade classes have the same bases either way."""
from collections import namedtuple

try:  # pragma: no cover - erlyx is not installable offline
    from erlyx import run_episodes
    from erlyx.callbacks import BaseCallback
    from erlyx.agents import BaseAgent, PolicyAgent
    from erlyx.environment import BaseEnvironment, Episode
    from erlyx.policies import Policy
    from erlyx.types import ActionData, EpisodeStatus
    HAVE_ERLYX = True
except Exception:  # noqa: BLE001
    HAVE_ERLYX = False
    ActionData = namedtuple('ActionData', ['action', 'info'])
    EpisodeStatus = namedtuple('EpisodeStatus', ['observation', 'reward', 'done'])

    class BaseAgent:
        def select_action(self, observation):
            raise NotImplementedError

    class PolicyAgent(BaseAgent):
        def __init__(self, policy):
            self._policy = policy

        @property
        def policy(self):
            return self._policy

    class Episode:
        pass

    class BaseEnvironment:
        def new_episode(self):
            raise NotImplementedError

    class Policy:
        pass

    class BaseCallback:
        def on_episode_begin(self, initial_observation):
            pass

        def on_step_end(self, action, observation, reward, done):
            pass

        def on_episode_end(self):
            pass

    def run_episodes(environment, agent, n_episodes, callbacks=None, use_tqdm=False):
        """The episode loop the reference drives (app/base.py:116-120), as inferred from its use: a truthy
        return from `on_episode_end` stops the remaining episodes (exp/callbacks.py:54, app/base.py:57,62)."""
        callbacks = list(callbacks or [])
        for _ in range(n_episodes):
            episode, observation = environment.new_episode()
            for cb in callbacks:
                cb.on_episode_begin(observation)
            done = False
            while not done:
                action = agent.select_action(observation)
                observation, reward, done = episode.step(action.action)
                for cb in callbacks:
                    cb.on_step_end(action, observation, reward, done)
            if any([cb.on_episode_end() for cb in callbacks]):
                break
