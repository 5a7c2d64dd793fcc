from .formats import RPInstance, RPObs  # noqa: F401
from .env import RPEnv  # noqa: F401
