"""A stand-in for the `gym` package, for ONE purpose: letting the unmodified reference worker
(rlgym_ppo/batched_agents/batched_agent.py:17-20, :185-196 imports gym and compares the action space's type with
gym.spaces.multi_discrete.MultiDiscrete / gym.spaces.box.Box) run in this container, where gym is not installed, inside
tests/golden/make_golden_wire.py.  Test infrastructure only; nothing in the product imports it."""
from . import spaces  # noqa: F401
