"""jampr_plus_b200: the B200 (sm_100a) construction-rollout hot path of JAMPR+ behind the reference's own
``RPEnv`` / ``Policy`` interfaces.  Importing the env or the policy loads libjampr_b200.so; there is no
CPU or PyTorch fallback."""
from .routing.formats import RPInstance, RPObs  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name == "RPEnv":
        from .routing.env import RPEnv
        return RPEnv
    if name in ("Policy", "GNN_ATTN_POLICY_CFG"):
        from .model import policy as _p
        return getattr(_p, name)
    if name in ("rollout", "rollout_sharded", "eval_episode"):
        from . import driver as _r
        return getattr(_r, name)
    raise AttributeError(name)
