def use(*a, **k):
    pass
from . import pyplot, animation, colors, cm  # noqa
