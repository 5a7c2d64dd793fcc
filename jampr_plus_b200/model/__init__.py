from .policy import Policy, GNN_ATTN_POLICY_CFG  # noqa: F401
