class PropertyNotImplementedError(Exception):
    pass


class SinglePointCalculator:
    pass
