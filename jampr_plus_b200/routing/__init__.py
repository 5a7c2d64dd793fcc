from .formats import RPInstance, RPObs  # noqa: F401
