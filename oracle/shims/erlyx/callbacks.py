class BaseCallback:
    def on_episode_begin(self, initial_observation):
        pass

    def on_step_end(self, action, observation, reward, done):
        pass

    def on_episode_end(self):
        pass
