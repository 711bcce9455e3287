def QLinearImplementation(*args, **kwargs):
    def deco(cls):
        return cls
    return deco
