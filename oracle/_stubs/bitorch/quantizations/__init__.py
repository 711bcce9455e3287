class Sign:
    pass


class SwishSign:
    pass
