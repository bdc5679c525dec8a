class _Constraint:
    def __init__(self, name):
        self.name = name


real_vector = _Constraint("real_vector")
real = _Constraint("real")
