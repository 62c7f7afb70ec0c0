"""Stand-in for `aenum` (absent in this image): the reference only needs Enum-like classes at import time."""
from enum import *  # noqa: F401,F403
from enum import Enum, unique  # noqa: F401


def extend_enum(*a, **k):
    raise NotImplementedError
