class Policy:
    def get_distribution(self, observation):
        raise NotImplementedError

    def num_actions(self):
        raise NotImplementedError
