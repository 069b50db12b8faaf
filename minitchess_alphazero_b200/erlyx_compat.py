"""The erlyx names the reference's self-play path imports (exp/agent.py:1-2, exp/environment.py:2,4,
exp/policy.py:1).  Uses the real `erlyx` when it is installed, otherwise minimal equivalents, so
the facade classes have the same bases either way."""
from collections import namedtuple

try:  # pragma: no cover - erlyx is not installable offline
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
