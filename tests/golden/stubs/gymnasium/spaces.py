class Discrete:
    def __init__(self, n):
        self.n = n


class Box:
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class Tuple(list):
    def __init__(self, spaces):
        super().__init__(spaces)
