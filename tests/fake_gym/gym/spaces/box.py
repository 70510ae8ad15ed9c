class Box:
    def __init__(self, shape):
        self.shape = tuple(shape)

    def seed(self, seed):
        pass
