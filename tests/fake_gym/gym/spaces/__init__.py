from . import box, multi_discrete  # noqa: F401


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()

    def seed(self, seed):
        pass
