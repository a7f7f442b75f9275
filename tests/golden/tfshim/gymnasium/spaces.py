class Space:
    shape = ()


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=None):
        self.low, self.high = low, high
        self.shape = tuple(int(s) for s in shape)
