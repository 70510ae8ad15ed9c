class MultiDiscrete:
    def __init__(self, nvec):
        self.nvec = list(nvec)
        self.shape = (len(self.nvec),)

    def seed(self, seed):
        pass
