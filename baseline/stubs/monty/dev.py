def requires(condition, message):
    """Identity decorator (the real one raises when an optional dependency is missing)."""
    return lambda obj: obj
