class BaseAgent:
    def select_action(self, observation):
        raise NotImplementedError


class PolicyAgent(BaseAgent):
    def __init__(self, policy):
        self._policy = policy

    @property
    def policy(self):
        return self._policy
