class BaseLearner:
    def update(self, dataset):
        raise NotImplementedError
