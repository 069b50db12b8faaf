class Episode:
    def step(self, action):
        raise NotImplementedError


class BaseEnvironment:
    def new_episode(self):
        raise NotImplementedError
