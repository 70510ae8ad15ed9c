"""rlgym_ppo_b200: B200-native (sm_100a) learner hot path with the rlgym-ppo surface.

`from rlgym_ppo_b200 import Learner` mirrors `from rlgym_ppo import Learner` (rlgym_ppo/__init__.py:1).
Importing the package needs the built C-ABI library (python -m rlgym_ppo_b200.build); there is no fallback.
"""
from . import _lib  # noqa: F401  (fails loudly if librlppo_b200.so is missing)

__version__ = "0.1.0"


def __getattr__(name):
    if name == "Learner":
        from .learner import Learner
        return Learner
    raise AttributeError(name)
